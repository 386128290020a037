#!/usr/bin/env python
"""Drop-in for the reference's contest_dilated_random.py (GRSS-DFC2014 visible scene) on the B200-native hot path.

Positional command line (contest:1228-1270):
  path output_path currentModelPath learningRate weight_decay batch_size niter crop_size stride_crop net_type
  distribution_type probValues update_type operation[train|test]
Scenes are read as ``<path>/train_image.npy`` / ``train_labels.npy`` / ``test_image.npy`` / ``test_labels.npy``
(float32 [H,W,3], uint8 [H,W] with 7 = unlabelled); see cli.load_npy_scenes for why.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import drs_b200  # noqa: E402,F401
from drs_b200 import cli, host, loops  # noqa: E402
from drs_b200.host import BatchColors  # noqa: E402

NUM_CLASSES = 7
NET_TYPES = ('dilated_icpr_old', 'dilated_grsl_old', 'dilated_grsl', 'dilated_icpr_rate6_densely', 'dilated_grsl_rate8', 'dilated8_grsl',
             'dilated_icpr_rate6', 'dilated_icpr_rate6_nodilation', 'dilated_icpr_rate6_SE', 'dilated_icpr_rate6_squeeze')


def main():
    list_params = ['path', 'output_path(for model, images, etc)', 'currentModelPath', 'learningRate', 'weight_decay',
                   'batch_size', 'niter', 'crop_size', 'stride_crop', 'net_type[' + '|'.join(NET_TYPES) + ']',
                   'distribution_type[single_fixed|multi_fixed|uniform|multinomial]', 'probValues', 'update_type [acc|loss]',
                   'operation [train|test]']
    if len(sys.argv) < len(list_params) + 1:
        sys.exit('Usage: ' + sys.argv[0] + ' ' + ' '.join(list_params))
    cli.print_params(list_params)
    a = sys.argv
    path, output_path, current_model = a[1], a[2], a[3]
    lr_initial, weight_decay, batch_size, niter = float(a[4]), float(a[5]), int(a[6]), int(a[7])
    crop_size, stride_crop, net_type, distribution_type = int(a[8]), int(a[9]), a[10], a[11]
    values = [int(i) for i in a[12].split(',')]
    update_type, operation = a[13], a[14]
    if net_type not in NET_TYPES:
        print(BatchColors.FAIL + 'Error! Net type not identified: ' + net_type + BatchColors.ENDC)
        return
    # contest starts patch_occur at ONES (contest:1275-1279)
    patch_acc_loss, patch_occur, patch_chosen_values = host.init_score_arrays(distribution_type, values, occur_init=1)
    probs = host.define_multinomial_probs(values) if distribution_type == 'multinomial' else None
    (training_data, test_data), (training_mask_data, test_mask_data) = cli.load_npy_scenes(path, ['train', 'test'])
    class_distribution = host.contest_create_distributions_over_classes(training_mask_data, crop_size, stride_crop, NUM_CLASSES)
    mean_full, std_full = host.contest_create_mean_and_std(training_data, class_distribution, crop_size)
    be = cli.make_backend(net_type, 3, NUM_CLASSES, weight_decay, lr_initial, 0.1, [training_data, test_data],
                          [training_mask_data, test_mask_data], mean_full, std_full, False, operation == 'train')
    if operation == 'train':
        loops.contest_train(be, training_data, training_mask_data, test_mask_data, class_distribution, output_path, current_model,
                            batch_size, niter, distribution_type, update_type, patch_acc_loss, patch_occur,
                            patch_chosen_values, probs, values, NUM_CLASSES)
    elif operation == 'test':
        be.restore(current_model)
        step = int(current_model.split('-')[-1]) if '-' in current_model else 0
        cur_val = int(values[0])
        if distribution_type in loops.DYNAMIC:
            pal = np.load(output_path + 'patch_acc_loss_step_' + str(step) + '.npy')
            occ = np.load(output_path + 'patch_occur_step_' + str(step) + '.npy')
            cur_val = host.select_best_patch_size(distribution_type, values, pal, occ, update_type, debug=True)
        loops.contest_test(be, test_mask_data, batch_size, step, cur_val, NUM_CLASSES)
    else:
        print(BatchColors.FAIL + "Process " + operation + "not found!" + BatchColors.ENDC)


if __name__ == "__main__":
    main()

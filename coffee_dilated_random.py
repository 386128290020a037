#!/usr/bin/env python
"""Drop-in for the reference's coffee_dilated_random.py (multispectral coffee tiles, two classes) on the B200-native path.

Positional command line (coffee:1105-1152), always trains:
  path_train path_test output_path currentModelPath learningRate weight_decay batch_size niter referenced_crop_size
  referenced_stride_crop net_type distribution_type probValues update_type
Tiles are read as ``<path>/tiles_image.npy`` ([N,H,W,3] float32) and ``tiles_labels.npy`` ([N,H,W] uint8).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import drs_b200  # noqa: E402,F401
from drs_b200 import cli, host, loops  # noqa: E402
from drs_b200.host import BatchColors  # noqa: E402

NUM_CLASSES = 2
NET_TYPES = ('dilated_icpr_original', 'dilated_grsl', 'dilated_icpr_rate6_densely', 'dilated_grsl_rate8', 'dilated8_grsl',
             'dilated_icpr_rate6', 'dilated_icpr_rate6_small', 'dilated_icpr_rate1', 'dilated_icpr_vary_rate', 'dilated_icpr_rate6_nodilation',
             'dilated_icpr_rate6_avgpool', 'dilated_icpr_rate6_SE', 'dilated_icpr_rate6_squeeze')


def load_tiles(path):
    img = np.load(os.path.join(path, 'tiles_image.npy')).astype(np.float32)
    lab = np.load(os.path.join(path, 'tiles_labels.npy')).astype(np.uint8)
    return [img[i] for i in range(len(img))], [lab[i].reshape(img.shape[1], img.shape[2]) for i in range(len(lab))]


def main():
    list_params = ['path_train', 'path_test', 'output_path(for model, images, etc)', 'currentModelPath', 'learningRate',
                   'weight_decay', 'batch_size', 'niter', 'referenced_crop_size', 'referenced_stride_crop',
                   'net_type[' + '|'.join(NET_TYPES) + ']', 'distribution_type[single_fixed|multi_fixed|uniform|multinomial]',
                   'probValues', 'update_type [acc|loss]']
    if len(sys.argv) < len(list_params) + 1:
        sys.exit('Usage: ' + sys.argv[0] + ' ' + ' '.join(list_params))
    cli.print_params(list_params)
    a = sys.argv
    path_train, path_test, output_path, current_model = a[1], a[2], a[3], a[4]
    lr_initial, weight_decay, batch_size, niter = float(a[5]), float(a[6]), int(a[7]), int(a[8])
    referenced_crop_size, referenced_stride_crop, net_type, distribution_type = int(a[9]), int(a[10]), a[11], a[12]
    values = [int(i) for i in a[13].split(',')]
    update_type = a[14]
    if net_type not in NET_TYPES:
        print(BatchColors.FAIL + 'Error! Net type not identified: ' + net_type + BatchColors.ENDC)
        return
    patch_acc_loss, patch_occur, patch_chosen_values = host.init_score_arrays(distribution_type, values)
    probs = host.define_multinomial_probs(values) if distribution_type == 'multinomial' else None
    training_data, training_mask_data = load_tiles(path_train)
    test_data, test_mask_data = load_tiles(path_test)
    class_distribution = host.coffee_create_distributions_over_classes(training_mask_data, referenced_crop_size,
                                                                       referenced_stride_crop, NUM_CLASSES)
    mean_full, std_full = host.coffee_create_mean_and_std(training_data, referenced_crop_size, referenced_stride_crop)
    be = cli.make_backend(net_type, 3, NUM_CLASSES, weight_decay, lr_initial, 0.1, training_data + test_data,
                          training_mask_data + test_mask_data, mean_full, std_full, False, True, train_fp16_patches=True)
    loops.coffee_train(be, training_data, test_mask_data, class_distribution, output_path, current_model, batch_size, niter,
                       distribution_type, update_type, patch_acc_loss, patch_occur, patch_chosen_values, probs, values,
                       NUM_CLASSES)


if __name__ == "__main__":
    main()

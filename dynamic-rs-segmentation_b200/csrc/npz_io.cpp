// Uncompressed .npz reader / writer (see npz_io.h).  Host only, no dependencies.
//
// File layout written here, member by member: ZIP local header (version 2.0, method 0 = stored, CRC-32, sizes), the
// member name "<variable>.npy", then the NPY 1.0 payload: "\x93NUMPY\x01\x00", a little-endian u16 header length, the
// dict "{'descr': '<f4', 'fortran_order': False, 'shape': (..), }" padded with spaces to a multiple of 64 bytes and
// closed by '\n', then the raw float32 data.  Central directory and end record follow.  numpy.savez itself writes ZIP64
// local headers (sizes 0xFFFFFFFF + an extra field); the reader therefore takes sizes and offsets from the central
// directory (and its ZIP64 extra field when present) and skips whatever the local header carries.
#include "npz_io.h"

#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unistd.h>

namespace drs_npz {

uint32_t crc32(const void* data, size_t n, uint32_t crc) {
  static uint32_t table[256];
  static bool init = false;
  if (!init) {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
      table[i] = c;
    }
    init = true;
  }
  const uint8_t* p = static_cast<const uint8_t*>(data);
  crc = ~crc;
  for (size_t i = 0; i < n; ++i) crc = table[(crc ^ p[i]) & 0xFF] ^ (crc >> 8);
  return ~crc;
}

namespace {

void put16(std::vector<uint8_t>& b, uint32_t v) { b.push_back(v & 0xFF); b.push_back((v >> 8) & 0xFF); }
void put32(std::vector<uint8_t>& b, uint32_t v) { put16(b, v & 0xFFFF); put16(b, v >> 16); }
uint32_t get16(const uint8_t* p) { return p[0] | (p[1] << 8); }
uint32_t get32(const uint8_t* p) { return get16(p) | (get16(p + 2) << 16); }
uint64_t get64(const uint8_t* p) { return (uint64_t)get32(p) | ((uint64_t)get32(p + 4) << 32); }

std::string npy_header(const std::vector<int64_t>& shape) {
  std::string d = "{'descr': '<f4', 'fortran_order': False, 'shape': (";
  for (size_t i = 0; i < shape.size(); ++i) {
    d += std::to_string((long long)shape[i]);
    if (shape.size() == 1 || i + 1 < shape.size()) d += ",";
    if (i + 1 < shape.size()) d += " ";
  }
  d += "), }";
  // magic (6) + version (2) + length (2) + dict + '\n' is a multiple of 64
  size_t total = 10 + d.size() + 1;
  d.append((64 - total % 64) % 64, ' ');
  d += "\n";
  std::string h("\x93NUMPY\x01\x00", 8);
  h.push_back((char)(d.size() & 0xFF));
  h.push_back((char)((d.size() >> 8) & 0xFF));
  return h + d;
}

std::string err(const std::string& path, const std::string& what) { return "checkpoint '" + path + "': " + what; }

// value of 'key' in the NPY header dict, e.g. "'<f4'", "False", "(3, 4)"
bool dict_value(const std::string& d, const std::string& key, std::string& out) {
  size_t k = d.find("'" + key + "'");
  if (k == std::string::npos) return false;
  size_t c = d.find(':', k);
  if (c == std::string::npos) return false;
  size_t b = d.find_first_not_of(" ", c + 1);
  if (b == std::string::npos) return false;
  size_t e;
  if (d[b] == '(') e = d.find(')', b);
  else if (d[b] == '\'') e = d.find('\'', b + 1);
  else { e = d.find_first_of(",}", b); if (e != std::string::npos) --e; }
  if (e == std::string::npos) return false;
  out = d.substr(b, e - b + 1);
  return true;
}

}  // namespace

std::string write(const std::string& path, const std::vector<Array>& arrays) {
  const std::string tmp = path + ".tmp." + std::to_string((long long)getpid());
  FILE* f = fopen(tmp.c_str(), "wb");
  if (!f) return err(path, std::string("cannot create ") + tmp + ": " + strerror(errno));
  std::vector<uint8_t> central;
  uint64_t offset = 0;
  auto fail = [&](const std::string& what) {
    fclose(f);
    remove(tmp.c_str());
    return err(path, what);
  };
  for (const Array& a : arrays) {
    if ((int64_t)a.data.size() != a.count()) return fail("array '" + a.name + "': shape does not match the element count");
    const std::string name = a.name + ".npy";
    const std::string hdr = npy_header(a.shape);
    const uint64_t size = hdr.size() + a.data.size() * 4;
    if (size >= 0xFFFFFFFFull || offset >= 0xFFFFFFFFull) return fail("array '" + a.name + "' needs ZIP64 (>= 4 GB), not supported");
    uint32_t crc = crc32(hdr.data(), hdr.size());
    crc = crc32(a.data.data(), a.data.size() * 4, crc);
    std::vector<uint8_t> lh;
    put32(lh, 0x04034b50); put16(lh, 20); put16(lh, 0); put16(lh, 0);       // signature, version, flags, method = stored
    put16(lh, 0); put16(lh, 0x21);                                           // time 00:00, date 1980-01-01
    put32(lh, crc); put32(lh, (uint32_t)size); put32(lh, (uint32_t)size);
    put16(lh, (uint32_t)name.size()); put16(lh, 0);
    if (fwrite(lh.data(), 1, lh.size(), f) != lh.size() || fwrite(name.data(), 1, name.size(), f) != name.size() ||
        fwrite(hdr.data(), 1, hdr.size(), f) != hdr.size() ||
        fwrite(a.data.data(), 4, a.data.size(), f) != a.data.size())
      return fail(std::string("write failed: ") + strerror(errno));
    put32(central, 0x02014b50); put16(central, 20); put16(central, 20); put16(central, 0); put16(central, 0);
    put16(central, 0); put16(central, 0x21);
    put32(central, crc); put32(central, (uint32_t)size); put32(central, (uint32_t)size);
    put16(central, (uint32_t)name.size()); put16(central, 0); put16(central, 0);   // name, extra, comment lengths
    put16(central, 0); put16(central, 0); put32(central, 0);                        // disk, internal, external attributes
    put32(central, (uint32_t)offset);
    central.insert(central.end(), name.begin(), name.end());
    offset += lh.size() + name.size() + size;
  }
  if (arrays.size() > 0xFFFE) return fail("too many arrays");
  std::vector<uint8_t> end;
  put32(end, 0x06054b50); put16(end, 0); put16(end, 0);
  put16(end, (uint32_t)arrays.size()); put16(end, (uint32_t)arrays.size());
  put32(end, (uint32_t)central.size()); put32(end, (uint32_t)offset); put16(end, 0);
  if (fwrite(central.data(), 1, central.size(), f) != central.size() || fwrite(end.data(), 1, end.size(), f) != end.size())
    return fail(std::string("write failed: ") + strerror(errno));
  if (fclose(f) != 0) { remove(tmp.c_str()); return err(path, std::string("close failed: ") + strerror(errno)); }
  if (rename(tmp.c_str(), path.c_str()) != 0) {
    const std::string m = strerror(errno);
    remove(tmp.c_str());
    return err(path, "rename failed: " + m);
  }
  return "";
}

std::string read(const std::string& path, std::vector<Array>& arrays) {
  arrays.clear();
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return err(path, std::string("cannot open: ") + strerror(errno));
  std::vector<uint8_t> buf;
  {
    fseek(f, 0, SEEK_END);
    const long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    if (n < 22) { fclose(f); return err(path, "not a zip archive (too short)"); }
    buf.resize((size_t)n);
    const size_t got = fread(buf.data(), 1, buf.size(), f);
    fclose(f);
    if (got != buf.size()) return err(path, "short read");
  }
  // end-of-central-directory record: last occurrence of its signature within the trailing 64 KB + 22 bytes
  size_t eocd = std::string::npos;
  for (size_t i = buf.size() - 22;; --i) {
    if (get32(&buf[i]) == 0x06054b50) { eocd = i; break; }
    if (i == 0 || buf.size() - i > 65536 + 22) break;
  }
  if (eocd == std::string::npos) return err(path, "not a zip archive (no end record)");
  uint64_t n_entries = get16(&buf[eocd + 10]);
  uint64_t cd_off = get32(&buf[eocd + 16]);
  if ((n_entries == 0xFFFF || cd_off == 0xFFFFFFFFu) && eocd >= 20 && get32(&buf[eocd - 20]) == 0x07064b50) {
    const uint64_t z64 = get64(&buf[eocd - 20 + 8]);       // ZIP64 end record
    if (buf.size() >= 56 && z64 <= buf.size() - 56 && get32(&buf[z64]) == 0x06064b50) {
      n_entries = get64(&buf[z64 + 32]);
      cd_off = get64(&buf[z64 + 48]);
    }
  }
  if (cd_off > buf.size()) return err(path, "corrupt end record");
  size_t p = (size_t)cd_off;
  for (uint64_t e = 0; e < n_entries; ++e) {
    if (p + 46 > buf.size() || get32(&buf[p]) != 0x02014b50) return err(path, "corrupt central directory");
    const uint32_t method = get16(&buf[p + 10]);
    const uint32_t crc = get32(&buf[p + 16]);
    uint64_t csize = get32(&buf[p + 20]), usize = get32(&buf[p + 24]);
    const uint32_t nlen = get16(&buf[p + 28]), xlen = get16(&buf[p + 30]), clen = get16(&buf[p + 32]);
    uint64_t lho = get32(&buf[p + 42]);
    if (p + 46 + nlen + xlen + clen > buf.size()) return err(path, "corrupt central directory");
    std::string name((const char*)&buf[p + 46], nlen);
    // ZIP64 extra field: the 64-bit values replace, in this order, whichever 32-bit fields are saturated
    size_t x = p + 46 + nlen;
    const size_t xend = x + xlen;
    while (x + 4 <= xend) {
      const uint32_t id = get16(&buf[x]), sz = get16(&buf[x + 2]);
      if (id == 0x0001) {
        size_t q = x + 4;
        if (usize == 0xFFFFFFFFu && q + 8 <= xend) { usize = get64(&buf[q]); q += 8; }
        if (csize == 0xFFFFFFFFu && q + 8 <= xend) { csize = get64(&buf[q]); q += 8; }
        if (lho == 0xFFFFFFFFu && q + 8 <= xend) { lho = get64(&buf[q]); q += 8; }
      }
      x += 4 + sz;
    }
    p += 46 + nlen + xlen + clen;
    if (method != 0) return err(path, "member '" + name + "' is compressed (numpy.savez_compressed): save with numpy.savez");
    if (csize != usize) return err(path, "member '" + name + "': stored sizes disagree");
    if (lho > buf.size() || lho + 30 > buf.size() || get32(&buf[lho]) != 0x04034b50) return err(path, "member '" + name + "': bad local header");
    const size_t data = (size_t)lho + 30 + get16(&buf[lho + 26]) + get16(&buf[lho + 28]);
    if (usize > buf.size() || data > buf.size() - usize) return err(path, "member '" + name + "' is truncated");
    if (crc32(&buf[data], (size_t)usize) != crc) return err(path, "member '" + name + "': CRC mismatch");
    if (name.size() < 4 || name.compare(name.size() - 4, 4, ".npy") != 0) continue;     // not an array: ignore
    // ---- NPY payload
    const uint8_t* d = &buf[data];
    if (usize < 10 || memcmp(d, "\x93NUMPY", 6) != 0) return err(path, "member '" + name + "' is not an NPY array");
    size_t hlen, hoff;
    if (d[6] == 1) { hlen = get16(d + 8); hoff = 10; }
    else if (d[6] == 2 || d[6] == 3) { if (usize < 12) return err(path, "member '" + name + "': short header"); hlen = get32(d + 8); hoff = 12; }
    else return err(path, "member '" + name + "': unknown NPY version");
    if (hoff + hlen > usize) return err(path, "member '" + name + "': short header");
    const std::string dict((const char*)d + hoff, hlen);
    std::string descr, fortran, shape;
    if (!dict_value(dict, "descr", descr) || !dict_value(dict, "fortran_order", fortran) || !dict_value(dict, "shape", shape))
      return err(path, "member '" + name + "': cannot parse the NPY header");
    if (fortran.find("False") == std::string::npos) return err(path, "member '" + name + "' is in Fortran order");
    Array a;
    a.name = name.substr(0, name.size() - 4);
    for (size_t i = 0; i < shape.size();) {
      if (shape[i] >= '0' && shape[i] <= '9') {
        char* endp = nullptr;
        a.shape.push_back(strtoll(shape.c_str() + i, &endp, 10));
        i = (size_t)(endp - shape.c_str());
      } else {
        ++i;
      }
    }
    const int64_t n = a.count();
    const uint8_t* payload = d + hoff + hlen;
    const uint64_t avail = usize - hoff - hlen;
    if (n < 0 || (uint64_t)n > avail) return err(path, "member '" + name + "': size does not match its shape");   // before allocating
    a.data.resize((size_t)n);
    auto need = [&](int es) { return (uint64_t)n * es == avail; };
    if (descr == "'<f4'") {
      if (!need(4)) return err(path, "member '" + name + "': size does not match its shape");
      memcpy(a.data.data(), payload, (size_t)n * 4);
    } else if (descr == "'<f8'") {
      if (!need(8)) return err(path, "member '" + name + "': size does not match its shape");
      for (int64_t i = 0; i < n; ++i) { double v; memcpy(&v, payload + i * 8, 8); a.data[i] = (float)v; }
    } else if (descr == "'<i4'") {
      if (!need(4)) return err(path, "member '" + name + "': size does not match its shape");
      for (int64_t i = 0; i < n; ++i) { int32_t v; memcpy(&v, payload + i * 4, 4); a.data[i] = (float)v; }
    } else if (descr == "'<i8'") {
      if (!need(8)) return err(path, "member '" + name + "': size does not match its shape");
      for (int64_t i = 0; i < n; ++i) { int64_t v; memcpy(&v, payload + i * 8, 8); a.data[i] = (float)v; }
    } else if (descr == "'|u1'") {
      if (!need(1)) return err(path, "member '" + name + "': size does not match its shape");
      for (int64_t i = 0; i < n; ++i) a.data[i] = (float)payload[i];
    } else {
      return err(path, "member '" + name + "': dtype " + descr + " is not supported (float32 expected)");
    }
    arrays.push_back(std::move(a));
  }
  return "";
}

}  // namespace drs_npz

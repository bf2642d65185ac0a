// host_io.inl - the reference's per-pair result files written from C (no GPU, no Python objects): included by roreg_capi.cu.
//
// The scene driver emits four files per pair: match_{keynum}/{id0}-{id1}.npy int64 [K,2] and scores/...npy float64
// [K] = ones (test/matcher.py:108-109), DR_index/...npy int64 [K] (test/estimator.py:111) and {yohoc}/{iters}iters/...npz with
// `trans` float64 [4,4] and `recalltime` int64 scalar (test/estimator.py:242).  Written through NumPy they cost ~0.4 ms of
// GIL-bound Python per pair and do not scale over threads (4 writer threads were slower than one); a ctypes call releases the
// GIL, so these plain-C writers run in parallel writer threads.  Formats: NPY 1.0 (magic, little-endian header length, dict padded
// with spaces to a multiple of 64, newline) and a ZIP archive of STORED members with CRC-32, which is what np.load reads.
#include <cstdio>
#include <cstring>
#include <cstdint>
#include <string>

namespace roreg_io {

static uint32_t crc32_update(uint32_t crc, const void* data, size_t n) {
  static uint32_t table[256]; static bool init = false;
  if (!init) {
    for (uint32_t i = 0; i < 256; ++i) { uint32_t c = i; for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1; table[i] = c; }
    init = true;
  }
  const uint8_t* p = static_cast<const uint8_t*>(data);
  crc = ~crc;
  for (size_t i = 0; i < n; ++i) crc = table[(crc ^ p[i]) & 0xFF] ^ (crc >> 8);
  return ~crc;
}

// NPY 1.0 header for a C-ordered array: descr e.g. "<i8", shape text e.g. "(3400, 2)", "(3400,)", "()"
static std::string npy_header(const char* descr, const std::string& shape) {
  std::string dict = std::string("{'descr': '") + descr + "', 'fortran_order': False, 'shape': " + shape + ", }";
  size_t total = 10 + dict.size() + 1;                       // magic(6) + version(2) + len(2) + dict + '\n'
  const size_t pad = (64 - total % 64) % 64;
  dict.append(pad, ' '); dict.push_back('\n');
  std::string h("\x93NUMPY\x01\x00", 8);
  const uint16_t len = (uint16_t)dict.size();
  h.push_back((char)(len & 0xFF)); h.push_back((char)(len >> 8));
  return h + dict;
}

static bool write_npy(const char* path, const char* descr, const std::string& shape, const void* data, size_t bytes) {
  FILE* f = fopen(path, "wb");
  if (!f) return false;
  const std::string h = npy_header(descr, shape);
  bool ok = fwrite(h.data(), 1, h.size(), f) == h.size() && (bytes == 0 || fwrite(data, 1, bytes, f) == bytes);
  ok = (fclose(f) == 0) && ok;
  return ok;
}

static void put16(std::string& s, uint16_t v) { s.push_back((char)(v & 0xFF)); s.push_back((char)(v >> 8)); }
static void put32(std::string& s, uint32_t v) { for (int i = 0; i < 4; ++i) s.push_back((char)((v >> (8 * i)) & 0xFF)); }

// ZIP archive with STORED members (name, payload) - local headers, central directory, end record
static bool write_zip(const char* path, const std::string names[], const std::string payloads[], int n) {
  std::string out, central;
  for (int i = 0; i < n; ++i) {
    const uint32_t crc = crc32_update(0, payloads[i].data(), payloads[i].size());
    const uint32_t size = (uint32_t)payloads[i].size(), offset = (uint32_t)out.size();
    put32(out, 0x04034b50u); put16(out, 20); put16(out, 0); put16(out, 0); put16(out, 0); put16(out, 0x21);   // version, flags, method 0, time, date (1980-01-01)
    put32(out, crc); put32(out, size); put32(out, size); put16(out, (uint16_t)names[i].size()); put16(out, 0);
    out += names[i]; out += payloads[i];
    put32(central, 0x02014b50u); put16(central, 20); put16(central, 20); put16(central, 0); put16(central, 0); put16(central, 0); put16(central, 0x21);
    put32(central, crc); put32(central, size); put32(central, size); put16(central, (uint16_t)names[i].size());
    put16(central, 0); put16(central, 0); put16(central, 0); put16(central, 0); put32(central, 0); put32(central, offset);
    central += names[i];
  }
  const uint32_t cd_off = (uint32_t)out.size(), cd_size = (uint32_t)central.size();
  out += central;
  put32(out, 0x06054b50u); put16(out, 0); put16(out, 0); put16(out, (uint16_t)n); put16(out, (uint16_t)n); put32(out, cd_size); put32(out, cd_off); put16(out, 0);
  FILE* f = fopen(path, "wb");
  if (!f) return false;
  bool ok = fwrite(out.data(), 1, out.size(), f) == out.size();
  ok = (fclose(f) == 0) && ok;
  return ok;
}

}  // namespace roreg_io

extern "C" int roreg_write_pair_files(const char* match_path, const char* scores_path, const char* dr_index_path, const char* npz_path,
                                      const int64_t* matches, const int64_t* dr_index, int K, const double* pose4x4, long long recalltime) {
  using namespace roreg_io;
  if (K < 0 || (K > 0 && (!matches || !dr_index))) return ROREG_ERR_ARG;
  const std::string k = std::to_string(K);
  if (match_path && !write_npy(match_path, "<i8", "(" + k + ", 2)", matches, (size_t)K * 16)) return ROREG_ERR_IO;
  if (scores_path) {                                            // np.ones(K): float64
    std::string ones((size_t)K * 8, '\0');
    const double one = 1.0;
    for (int i = 0; i < K; ++i) memcpy(&ones[(size_t)i * 8], &one, 8);
    if (!write_npy(scores_path, "<f8", "(" + k + ",)", ones.data(), ones.size())) return ROREG_ERR_IO;
  }
  if (dr_index_path && !write_npy(dr_index_path, "<i8", "(" + k + ",)", dr_index, (size_t)K * 8)) return ROREG_ERR_IO;
  if (npz_path) {
    if (!pose4x4) return ROREG_ERR_ARG;
    std::string names[2] = {"trans.npy", "recalltime.npy"}, payloads[2];
    payloads[0] = npy_header("<f8", "(4, 4)"); payloads[0].append(reinterpret_cast<const char*>(pose4x4), 128);
    const int64_t r = (int64_t)recalltime;
    payloads[1] = npy_header("<i8", "()"); payloads[1].append(reinterpret_cast<const char*>(&r), 8);
    if (!write_zip(npz_path, names, payloads, 2)) return ROREG_ERR_IO;
  }
  return ROREG_OK;
}

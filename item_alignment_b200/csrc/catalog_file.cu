// Binary catalog files: the [C, D] embedding matrix + an item-id table, so that retrieval over 10^8 rows
// starts from one mmap instead of parsing text.  Replaces the embedding JSONL the reference writes
// (finetune_text.py:784-792: one pair per line, each embedding a stringified float list) and reads back with `eval`
// (model_ensemble.py:112).  Host-side C++ only; the one device interaction is the chunked upload.
//
// Layout (little-endian):
//   [0, 64)                     header: magic "IACATLG1", version, dtype, rows, dim, data_offset, ids_offset, ids_bytes
//   [data_offset, +rows*dim*e)  row-major matrix, data_offset = 4096 (page aligned: the mmap'ed rows are 16-byte aligned)
//   [ids_offset, +ids_bytes)    optional: uint64 offsets[rows + 1], then the UTF-8 ids back to back
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cerrno>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "common.cuh"

namespace ia {
namespace {

constexpr char kMagic[8] = {'I', 'A', 'C', 'A', 'T', 'L', 'G', '1'};
constexpr uint64_t kDataOffset = 4096;

struct FileHeader {
  char magic[8];
  uint32_t version;
  uint32_t dtype;
  uint64_t rows, dim;
  uint64_t data_offset, ids_offset, ids_bytes;
  uint64_t reserved;
};
static_assert(sizeof(FileHeader) == 64, "header is 64 bytes");

size_t elem_bytes(int dtype) { return dtype == IA_F32 ? 4 : 2; }

// float -> bf16 / fp16 bits, round to nearest even (what torch's .to(dtype) does)
uint16_t f32_to_bf16(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);   // NaN stays NaN
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
uint16_t f32_to_f16(float f) {
  const __half h = __float2half_rn(f);   // host path of cuda_fp16.h: IEEE round to nearest even
  uint16_t b;
  memcpy(&b, &h, 2);
  return b;
}

int write_all(int fd, const void* buf, size_t n) {
  const char* p = static_cast<const char*>(buf);
  while (n > 0) {
    const ssize_t w = write(fd, p, n);
    if (w < 0) { if (errno == EINTR) continue; return -1; }
    p += w; n -= (size_t)w;
  }
  return 0;
}

// Streams a catalog file: header placeholder, rows appended one at a time, ids gathered in memory, finish() patches the header.
struct Writer {
  int fd = -1;
  int dtype = IA_BF16;
  uint64_t rows = 0, dim = 0;
  std::vector<uint64_t> id_off;
  std::string id_blob;
  std::vector<uint8_t> rowbuf;
  bool with_ids = false;

  int open_file(const char* path, int dt, bool ids) {
    fd = ::open(path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (fd < 0) { set_error("cannot create %s: %s", path, strerror(errno)); return IA_ERR_INVALID; }
    dtype = dt; with_ids = ids;
    std::vector<uint8_t> zero(kDataOffset, 0);
    if (write_all(fd, zero.data(), zero.size()) != 0) { set_error("write failed: %s", strerror(errno)); return IA_ERR_CUDA; }
    id_off.push_back(0);
    return IA_OK;
  }
  int add_row(const float* v, uint64_t d, const char* id, size_t id_len) {
    if (rows == 0) { dim = d; rowbuf.resize(d * elem_bytes(dtype)); }
    if (d != dim) { set_error("row %llu has %llu values, expected %llu", (unsigned long long)rows, (unsigned long long)d, (unsigned long long)dim); return IA_ERR_INVALID; }
    if (dtype == IA_F32) memcpy(rowbuf.data(), v, d * 4);
    else {
      uint16_t* o = reinterpret_cast<uint16_t*>(rowbuf.data());
      for (uint64_t i = 0; i < d; ++i) o[i] = dtype == IA_BF16 ? f32_to_bf16(v[i]) : f32_to_f16(v[i]);
    }
    if (write_all(fd, rowbuf.data(), rowbuf.size()) != 0) { set_error("write failed: %s", strerror(errno)); return IA_ERR_CUDA; }
    if (with_ids) { id_blob.append(id, id_len); id_off.push_back(id_blob.size()); }
    ++rows;
    return IA_OK;
  }
  int add_raw_rows(const void* data, uint64_t n, uint64_t d) {   // already in the file dtype
    if (rows == 0) dim = d;
    if (write_all(fd, data, n * d * elem_bytes(dtype)) != 0) { set_error("write failed: %s", strerror(errno)); return IA_ERR_CUDA; }
    rows += n;
    return IA_OK;
  }
  int finish() {
    FileHeader h{};
    memcpy(h.magic, kMagic, 8);
    h.version = 1; h.dtype = (uint32_t)dtype; h.rows = rows; h.dim = dim; h.data_offset = kDataOffset;
    const uint64_t data_end = kDataOffset + rows * dim * elem_bytes(dtype);
    if (with_ids) {
      const uint64_t pad = (8 - data_end % 8) % 8;
      const char zeros[8] = {0};
      if (pad && write_all(fd, zeros, pad) != 0) { set_error("write failed: %s", strerror(errno)); return IA_ERR_CUDA; }
      h.ids_offset = data_end + pad;
      h.ids_bytes = id_off.size() * 8 + id_blob.size();
      if (write_all(fd, id_off.data(), id_off.size() * 8) != 0 || write_all(fd, id_blob.data(), id_blob.size()) != 0) {
        set_error("write failed: %s", strerror(errno));
        return IA_ERR_CUDA;
      }
    }
    if (pwrite(fd, &h, sizeof(h), 0) != (ssize_t)sizeof(h)) { set_error("header write failed: %s", strerror(errno)); return IA_ERR_CUDA; }
    return IA_OK;
  }
  ~Writer() { if (fd >= 0) ::close(fd); }
};

// value of a JSON string member `"key": "..."` on one line (the reference writes flat objects with json.dumps, so a key
// cannot occur inside another value except as escaped text, which the leading quote test excludes)
bool find_string_member(const char* line, size_t len, const char* key, const char** val, size_t* val_len) {
  const size_t klen = strlen(key);
  for (size_t i = 0; i + klen + 2 < len; ++i) {
    if (line[i] != '"' || (i > 0 && line[i - 1] == '\\')) continue;
    if (memcmp(line + i + 1, key, klen) != 0 || line[i + 1 + klen] != '"') continue;
    size_t j = i + klen + 2;
    while (j < len && (line[j] == ' ' || line[j] == ':')) ++j;
    if (j >= len || line[j] != '"') return false;
    const size_t start = ++j;
    while (j < len && !(line[j] == '"' && line[j - 1] != '\\')) ++j;
    if (j >= len) return false;
    *val = line + start; *val_len = j - start;
    return true;
  }
  return false;
}

// One decimal number -> the double Python's float() / eval would produce, without strtod's cost (Clinger's fast path):
// a decimal whose digit string is <= 2^53 with |exponent| <= 22 is the quotient / product of two exact doubles, i.e. the
// correctly rounded double.  Everything else (nan, inf, > 17 digits, large exponents) goes to strtod.  The caller
// casts to float32 -- the same decimal -> float64 -> float32 route a consumer of the reference's file takes
// (eval at model_ensemble.py:112, then numpy's float32), so the values agree bit for bit by construction.
// Advances *pp past the number on success.
bool fast_decimal_to_double(const char** pp, double* out) {
  static const double kPow10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                    1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
  const char* s = *pp;
  bool neg = false;
  if (*s == '-') { neg = true; ++s; } else if (*s == '+') ++s;
  uint64_t m = 0;
  int sig = 0, exp10 = 0;
  bool any = false;
  while (*s >= '0' && *s <= '9') {
    if (sig < 18) { m = m * 10 + (uint64_t)(*s - '0'); if (m) ++sig; } else return false;
    ++s; any = true;
  }
  if (*s == '.') {
    ++s;
    while (*s >= '0' && *s <= '9') {
      if (sig < 18) { m = m * 10 + (uint64_t)(*s - '0'); if (m) ++sig; --exp10; } else return false;
      ++s; any = true;
    }
  }
  if (!any) return false;
  if (*s == 'e' || *s == 'E') {
    const char* e = s + 1;
    bool eneg = false;
    if (*e == '-') { eneg = true; ++e; } else if (*e == '+') ++e;
    if (!(*e >= '0' && *e <= '9')) return false;
    int ev = 0;
    while (*e >= '0' && *e <= '9') { if (ev < 10000) ev = ev * 10 + (*e - '0'); ++e; }
    exp10 += eneg ? -ev : ev;
    s = e;
  }
  if ((*s >= 'a' && *s <= 'z') || (*s >= 'A' && *s <= 'Z') || *s == '.') return false;   // nan, inf, hex floats, "1.2.3"
  if (m == 0) { *out = neg ? -0.0 : 0.0; *pp = s; return true; }
  if (m > (1ull << 53) || exp10 < -22 || exp10 > 22) return false;
  const double q = exp10 < 0 ? (double)m / kPow10[-exp10] : (double)m * kPow10[exp10];
  *out = neg ? -q : q;
  *pp = s;
  return true;
}

// "[0.1,-2.5e-3, ...]" -> floats: float32(float64(decimal)) for every value.  The float32 the reference printed (repr of a
// numpy float32 = shortest round-trip decimal) comes back bit for bit; a float64 repr (the ensemble's averaged scores) is
// rounded exactly like numpy rounds the evaluated Python float.
bool parse_float_list(const char* s, size_t len, std::vector<float>* out) {
  out->clear();
  std::string tmp(s, len);   // NUL-terminated copy
  const char* p = tmp.c_str();
  while (*p == ' ' || *p == '[') ++p;
  while (*p && *p != ']') {
    double v;
    if (!fast_decimal_to_double(&p, &v)) {
      char* end = nullptr;
      v = strtod(p, &end);   // correctly rounded double for any input
      if (end == p) return false;
      p = end;
    }
    out->push_back((float)v);
    while (*p == ' ' || *p == ',') ++p;
  }
  return true;
}

}  // namespace
}  // namespace ia

using namespace ia;

struct ia_catalog_file {
  int fd;
  void* map;
  size_t map_bytes;
  FileHeader hdr;
  const uint64_t* id_off;
  const char* id_blob;
};

extern "C" {

int ia_catalog_file_write(const char* path, int dtype, const void* data, int64_t rows, int64_t dim, const char* const* ids) {
  if (path == nullptr || (rows > 0 && data == nullptr) || rows < 0 || dim <= 0) { set_error("catalog file: bad arguments"); return IA_ERR_INVALID; }
  if (dtype != IA_F32 && dtype != IA_BF16 && dtype != IA_F16) { set_error("catalog file: unsupported dtype %d", dtype); return IA_ERR_UNSUPPORTED; }
  Writer w;
  int rc = w.open_file(path, dtype, ids != nullptr);
  if (rc != IA_OK) return rc;
  w.dim = (uint64_t)dim;
  if ((rc = w.add_raw_rows(data, (uint64_t)rows, (uint64_t)dim)) != IA_OK) return rc;
  if (ids != nullptr)
    for (int64_t i = 0; i < rows; ++i) { w.id_blob.append(ids[i] ? ids[i] : ""); w.id_off.push_back(w.id_blob.size()); }
  return w.finish();
}

int ia_catalog_file_open(const char* path, ia_catalog_file** out) {
  if (path == nullptr || out == nullptr) { set_error("catalog file: bad arguments"); return IA_ERR_INVALID; }
  const int fd = ::open(path, O_RDONLY);
  if (fd < 0) { set_error("cannot open %s: %s", path, strerror(errno)); return IA_ERR_INVALID; }
  struct stat st;
  if (fstat(fd, &st) != 0 || (size_t)st.st_size < sizeof(FileHeader)) { ::close(fd); set_error("%s: not a catalog file (too short)", path); return IA_ERR_INVALID; }
  void* map = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_SHARED, fd, 0);
  if (map == MAP_FAILED) { ::close(fd); set_error("mmap of %s failed: %s", path, strerror(errno)); return IA_ERR_CUDA; }
  ia_catalog_file* f = new ia_catalog_file{fd, map, (size_t)st.st_size, {}, nullptr, nullptr};
  memcpy(&f->hdr, map, sizeof(FileHeader));
  const FileHeader& h = f->hdr;
  const bool dtype_ok = h.dtype == IA_F32 || h.dtype == IA_BF16 || h.dtype == IA_F16;
  // every size below comes from the file: overflow-checked arithmetic, then bounds against the mapped size
  const uint64_t fsize = (uint64_t)st.st_size;
  uint64_t data_bytes = 0, data_end = 0, elems = 0;
  bool ok = memcmp(h.magic, kMagic, 8) == 0 && h.version == 1 && dtype_ok && h.dim > 0 && h.rows <= (uint64_t)INT64_MAX &&
            h.dim <= (uint64_t)INT64_MAX && h.data_offset >= sizeof(FileHeader) && h.data_offset % 16 == 0 &&
            !__builtin_mul_overflow(h.rows, h.dim, &elems) &&
            !__builtin_mul_overflow(elems, (uint64_t)elem_bytes((int)h.dtype), &data_bytes) &&
            !__builtin_add_overflow(h.data_offset, data_bytes, &data_end) && data_end <= fsize;
  if (ok && h.ids_offset != 0) {
    uint64_t ids_end = 0, table_bytes = 0;
    ok = h.ids_offset % 8 == 0 && h.ids_offset >= data_end && !__builtin_add_overflow(h.ids_offset, h.ids_bytes, &ids_end) &&
         ids_end <= fsize && !__builtin_mul_overflow(h.rows + 1, (uint64_t)8, &table_bytes) && h.ids_bytes >= table_bytes;
    if (ok) {
      f->id_off = reinterpret_cast<const uint64_t*>(static_cast<const char*>(map) + h.ids_offset);
      f->id_blob = reinterpret_cast<const char*>(f->id_off + h.rows + 1);
      const uint64_t blob_bytes = h.ids_bytes - table_bytes;
      ok = f->id_off[0] == 0 && f->id_off[h.rows] == blob_bytes;
      // the offset table must be non-decreasing (with the end fixed above that bounds every entry by the blob size):
      // ia_catalog_file_id hands out blob + off[row] with length off[row+1] - off[row]
      for (uint64_t i = 0; ok && i < h.rows; ++i) ok = f->id_off[i] <= f->id_off[i + 1];
    }
  }
  if (!ok) {
    munmap(map, (size_t)st.st_size); ::close(fd); delete f;
    set_error("%s: not a valid catalog file (bad magic, version, dtype or section bounds)", path);
    return IA_ERR_INVALID;
  }
  *out = f;
  return IA_OK;
}

void ia_catalog_file_close(ia_catalog_file* f) {
  if (f == nullptr) return;
  munmap(f->map, f->map_bytes);
  ::close(f->fd);
  delete f;
}

int ia_catalog_file_info(const ia_catalog_file* f, int* dtype, int64_t* rows, int64_t* dim, int* has_ids) {
  if (f == nullptr) { set_error("catalog file: null handle"); return IA_ERR_INVALID; }
  if (dtype) *dtype = (int)f->hdr.dtype;
  if (rows) *rows = (int64_t)f->hdr.rows;
  if (dim) *dim = (int64_t)f->hdr.dim;
  if (has_ids) *has_ids = f->id_off != nullptr;
  return IA_OK;
}

const void* ia_catalog_file_data(const ia_catalog_file* f) {
  return f == nullptr ? nullptr : static_cast<const char*>(f->map) + f->hdr.data_offset;
}

const char* ia_catalog_file_id(const ia_catalog_file* f, int64_t row, int64_t* len) {
  if (f == nullptr || f->id_off == nullptr || row < 0 || (uint64_t)row >= f->hdr.rows) { if (len) *len = 0; return nullptr; }
  if (len) *len = (int64_t)(f->id_off[row + 1] - f->id_off[row]);
  return f->id_blob + f->id_off[row];
}

// rows [row_begin, row_end) -> device memory (row-major, ld = dim), through two pinned staging buffers so that the
// page-cache read of chunk i+1 overlaps the DMA of chunk i.  Synchronises the stream before returning.
int ia_catalog_file_upload(const ia_catalog_file* f, int64_t row_begin, int64_t row_end, void* device_dst, ia_stream_t stream) {
  if (f == nullptr || device_dst == nullptr || row_begin < 0 || row_end < row_begin || (uint64_t)row_end > f->hdr.rows) {
    set_error("catalog upload: bad arguments");
    return IA_ERR_INVALID;
  }
  cudaStream_t s = (cudaStream_t)stream;
  const size_t row_bytes = f->hdr.dim * elem_bytes((int)f->hdr.dtype);
  const size_t total = (size_t)(row_end - row_begin) * row_bytes;
  const char* src = static_cast<const char*>(ia_catalog_file_data(f)) + (size_t)row_begin * row_bytes;
  const size_t chunk = 32u << 20;
  void* stage[2] = {nullptr, nullptr};
  cudaEvent_t done[2];
  IA_CUDA_CHECK(cudaMallocHost(&stage[0], chunk));
  if (cudaMallocHost(&stage[1], chunk) != cudaSuccess) { cudaFreeHost(stage[0]); set_error("pinned staging allocation failed"); return IA_ERR_CUDA; }
  cudaEventCreateWithFlags(&done[0], cudaEventDisableTiming);
  cudaEventCreateWithFlags(&done[1], cudaEventDisableTiming);
  int rc = IA_OK;
  int b = 0;
  for (size_t off = 0; off < total; off += chunk, b ^= 1) {
    const size_t n = total - off < chunk ? total - off : chunk;
    if (off >= 2 * chunk && cudaEventSynchronize(done[b]) != cudaSuccess) { rc = IA_ERR_CUDA; break; }   // staging buffer free again
    memcpy(stage[b], src + off, n);
    if (cudaMemcpyAsync(static_cast<char*>(device_dst) + off, stage[b], n, cudaMemcpyHostToDevice, s) != cudaSuccess ||
        cudaEventRecord(done[b], s) != cudaSuccess) { rc = IA_ERR_CUDA; break; }
  }
  if (cudaStreamSynchronize(s) != cudaSuccess) rc = IA_ERR_CUDA;
  if (rc != IA_OK) set_error("catalog upload failed: %s", cudaGetErrorString(cudaGetLastError()));
  cudaEventDestroy(done[0]); cudaEventDestroy(done[1]);
  cudaFreeHost(stage[0]); cudaFreeHost(stage[1]);
  return rc;
}

// Reference embedding JSONL (finetune_text.py:784-792) -> catalog file.  side: 0 = src_item_*, 1 = tgt_item_*, 2 = both.
// An item that occurs in several pairs is stored once (first occurrence wins, like a dict built in file order).
// The file is mmap'ed and parsed in blocks by `threads` workers (0 = all hardware threads), each on a run of whole lines;
// the blocks' records are then appended in file order by one thread, so the output does not depend on the thread count.
int ia_embedding_jsonl_to_catalog(const char* jsonl_path, const char* out_path, int dtype, int side, int threads,
                                  int64_t* rows_out, int64_t* dim_out) {
  if (jsonl_path == nullptr || out_path == nullptr || side < 0 || side > 2 || threads < 0) { set_error("jsonl: bad arguments"); return IA_ERR_INVALID; }
  if (dtype != IA_F32 && dtype != IA_BF16 && dtype != IA_F16) { set_error("jsonl: unsupported dtype %d", dtype); return IA_ERR_UNSUPPORTED; }
  const int fd = ::open(jsonl_path, O_RDONLY);
  if (fd < 0) { set_error("cannot open %s: %s", jsonl_path, strerror(errno)); return IA_ERR_INVALID; }
  struct stat st;
  if (fstat(fd, &st) != 0) { ::close(fd); set_error("cannot stat %s: %s", jsonl_path, strerror(errno)); return IA_ERR_INVALID; }
  const size_t size = (size_t)st.st_size;
  const char* data = nullptr;
  if (size > 0) {
    void* map = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
    if (map == MAP_FAILED) { ::close(fd); set_error("mmap of %s failed: %s", jsonl_path, strerror(errno)); return IA_ERR_CUDA; }
    data = static_cast<const char*>(map);
    madvise(map, size, MADV_SEQUENTIAL);
  }
  struct Unmap {
    const char* d; size_t n; int fd;
    ~Unmap() { if (d) munmap(const_cast<char*>(d), n); ::close(fd); }
  } unmap{data, size, fd};

  Writer w;
  int rc = w.open_file(out_path, dtype, true);
  if (rc != IA_OK) return rc;
  int nthreads = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 64) nthreads = 64;

  struct Chunk {   // what one worker extracts from its run of lines, in line order
    std::string ids;                 // ids back to back
    std::vector<uint32_t> id_len;    // one entry per record
    std::vector<float> vals;         // dim floats per record
    std::vector<uint32_t> dims;      // values in each record (all equal in a well-formed file)
    int64_t lines = 0;               // lines consumed (for error messages)
    int64_t err_line = -1;           // first bad line within the chunk, 1-based
    std::string err;
  };
  static const char* kId[2] = {"src_item_id", "tgt_item_id"};
  static const char* kEmb[2] = {"src_item_emb", "tgt_item_emb"};
  auto parse_range = [&](const char* p, const char* end, Chunk* c) {
    std::vector<float> vals;
    while (p < end) {
      const char* nl = static_cast<const char*>(memchr(p, '\n', (size_t)(end - p)));
      const char* le = nl ? nl : end;
      const size_t len = (size_t)(le - p);
      ++c->lines;
      bool blank = true;
      for (size_t i = 0; i < len; ++i) if (p[i] != ' ' && p[i] != '\r' && p[i] != '\t') { blank = false; break; }
      if (!blank) {
        for (int sd = 0; sd < 2; ++sd) {
          if (side != 2 && side != sd) continue;
          const char *id, *emb;
          size_t idl, embl;
          if (!find_string_member(p, len, kId[sd], &id, &idl) || !find_string_member(p, len, kEmb[sd], &emb, &embl)) {
            c->err_line = c->lines; c->err = std::string("missing ") + kId[sd] + " / " + kEmb[sd];
            return;
          }
          if (!parse_float_list(emb, embl, &vals) || vals.empty()) {
            c->err_line = c->lines; c->err = std::string(kEmb[sd]) + " is not a float list";
            return;
          }
          c->ids.append(id, idl);
          c->id_len.push_back((uint32_t)idl);
          c->vals.insert(c->vals.end(), vals.begin(), vals.end());
          c->dims.push_back((uint32_t)vals.size());
        }
      }
      p = nl ? nl + 1 : end;
    }
  };

  std::unordered_map<std::string, uint64_t> seen;
  // bytes of text per worker and block: an even share of the file, at most 32 MiB (bounds the memory held before the merge)
  size_t per_worker = (size + (size_t)nthreads - 1) / (size_t)nthreads;
  if (per_worker < (64u << 10)) per_worker = 64u << 10;
  if (per_worker > (32u << 20)) per_worker = 32u << 20;
  size_t pos = 0;
  int64_t lines_before = 0;
  while (pos < size && rc == IA_OK) {
    // cut [pos, block_end) into nthreads runs of whole lines
    std::vector<const char*> cuts;
    cuts.push_back(data + pos);
    for (int t = 1; t <= nthreads; ++t) {
      size_t target = pos + (size_t)t * per_worker;
      const char* cut = data + size;
      if (target < size) {
        const char* nl = static_cast<const char*>(memchr(data + target, '\n', size - target));
        cut = nl ? nl + 1 : data + size;
      }
      if (cut < cuts.back()) cut = cuts.back();
      cuts.push_back(cut);
      if (cut == data + size) break;
    }
    const int nchunks = (int)cuts.size() - 1;
    std::vector<Chunk> chunks((size_t)nchunks);
    std::vector<std::thread> pool;
    for (int t = 1; t < nchunks; ++t) pool.emplace_back(parse_range, cuts[t], cuts[t + 1], &chunks[t]);
    parse_range(cuts[0], cuts[1], &chunks[0]);
    for (auto& th : pool) th.join();
    // append in file order; the first error in file order wins
    for (int t = 0; t < nchunks && rc == IA_OK; ++t) {
      Chunk& c = chunks[t];
      size_t id_pos = 0, val_pos = 0;
      for (size_t r = 0; r < c.id_len.size() && rc == IA_OK; ++r) {
        std::string key(c.ids.data() + id_pos, c.id_len[r]);
        if (!seen.count(key)) {
          rc = w.add_row(c.vals.data() + val_pos, c.dims[r], key.data(), key.size());
          seen.emplace(std::move(key), w.rows);
        }
        id_pos += c.id_len[r];
        val_pos += c.dims[r];
      }
      if (rc == IA_OK && c.err_line >= 0) {
        set_error("%s:%lld: %s", jsonl_path, (long long)(lines_before + c.err_line), c.err.c_str());
        rc = IA_ERR_INVALID;
      }
      lines_before += c.lines;
    }
    pos = (size_t)(cuts.back() - data);
  }
  if (rc != IA_OK) return rc;
  if ((rc = w.finish()) != IA_OK) return rc;
  if (rows_out) *rows_out = (int64_t)w.rows;
  if (dim_out) *dim_out = (int64_t)w.dim;
  return IA_OK;
}

}  // extern "C"

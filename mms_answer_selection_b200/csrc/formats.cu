// Input formats of the path, host side: the pre-trained word-vector files that EmbedLayer::LayerSetUp reads when
// embed_param.weight_source is set (reference src/caffe/layers/embed_layer.cpp:46-113).  Host code only (no kernel):
// the reference fills blobs_[0]->mutable_cpu_data() during LayerSetUp and so does this; the table reaches the device
// through the blob's ordinary host->device sync.
//
// Three formats, chosen by the last three characters of the file name exactly as the reference does (:50-51, :62):
//   "...txt"  GloVe text            word v_0 ... v_{D-1}                        one word per record
//   "...all"  the fork's dump       header "<float> <V-1> <D-1>", then          <int id> v_0 ... v_{D-1} <word>
//   other     word2vec binary       header "<vocab> <dim>", then                word ' ' + dim raw little-endian floats
// Records fill the table from row 0 in file order; rows the file does not reach keep what the filler put there.
//
// Kept from the reference on purpose: every value is parsed/read as a 32-bit float and stored through a float* into
// the slot of element w (`(float*)(weight_data + w_index)`, :56, :74, :97) -- for double blobs that overwrites the
// first four bytes of the double and leaves the rest, which is what the reference's double instantiation does, so
// the resulting table is bit-identical in both instantiations.
// Different from the reference on purpose: the reference never checks w_index against the table size (a longer file
// overruns the blob) nor the fopen result; here a missing file, a record past the last row and a malformed record
// are errors (MMS_E_INVALID with a message).
#include <cerrno>
#include <cstdio>
#include <cstring>
#include <string>

#include "mms_common.cuh"

namespace {

struct File {
  FILE* f;
  explicit File(const char* path, const char* mode) : f(fopen(path, mode)) {}
  ~File() { if (f) fclose(f); }
};

template <typename T>
struct Table {
  T* data; long long rows, dim, filled = 0;               // filled = elements written so far
  bool full() const { return filled >= rows * dim; }
  void put(float v) { memcpy(reinterpret_cast<char*>(data + filled), &v, sizeof(float)); ++filled; }
};

bool ends_with(const std::string& s, const char* suffix) {
  const size_t n = strlen(suffix);
  return s.size() >= n && s.compare(s.size() - n, n, suffix) == 0;
}

template <typename T>
int read_floats_text(FILE* f, Table<T>& t, const char* path) {
  for (long long i = 0; i < t.dim; ++i) {
    float v;
    if (fscanf(f, "%f ", &v) != 1) {
      mms_set_error("weight_source %s: record %lld: expected %lld values, found %lld", path, t.filled / t.dim, t.dim, i);
      return MMS_E_INVALID;
    }
    t.put(v);
  }
  return 0;
}

#define OVERRUN_CHECK()                                                                                         \
  if (t.full()) {                                                                                               \
    mms_set_error("weight_source %s: more than input_dim = %lld records (the reference would overrun the blob)", \
                  path, t.rows);                                                                                \
    return MMS_E_INVALID;                                                                                       \
  }

template <typename T>
int load_glove_txt(FILE* f, Table<T>& t, const char* path) {
  char word[4096];
  while (fscanf(f, "%4095s ", word) == 1) {                 // :53
    OVERRUN_CHECK();
    MMS_TRY(read_floats_text(f, t, path));
  }
  return 0;
}

template <typename T>
int load_all(FILE* f, Table<T>& t, const char* path) {
  float b1;
  int t1, t2;
  if (fscanf(f, "%f %d %d", &b1, &t1, &t2) != 3) {           // :66
    mms_set_error("weight_source %s: bad header (expected '<float> <int> <int>')", path);
    return MMS_E_INVALID;
  }
  if (t1 != t.rows - 1 || t2 != t.dim - 1) {                 // CHECK_EQ(t1, K_-1); CHECK_EQ(t2, N_-1)  :67-68
    mms_set_error("weight_source %s: header says (%d, %d), the layer needs (input_dim - 1, num_output - 1) = (%lld, %lld)",
                  path, t1, t2, t.rows - 1, t.dim - 1);
    return MMS_E_INVALID;
  }
  char word[4096];
  for (;;) {
    const int got = fscanf(f, "%d ", &t1);                   // :69
    if (got == EOF) break;
    if (got != 1) {
      mms_set_error("weight_source %s: record %lld does not start with an integer id", path, t.filled / t.dim);
      return MMS_E_INVALID;
    }
    OVERRUN_CHECK();
    MMS_TRY(read_floats_text(f, t, path));
    if (fscanf(f, "%4095s", word) != 1) word[0] = 0;         // :74 (the trailing word; a missing one ends the file)
  }
  return 0;
}

template <typename T>
int load_word2vec_bin(FILE* f, Table<T>& t, const char* path) {
  long long vocab = 0, dim = 0;
  if (fscanf(f, "%lld", &vocab) != 1 || fscanf(f, "%lld", &dim) != 1) {   // :82-83
    mms_set_error("weight_source %s: bad word2vec header", path);
    return MMS_E_INVALID;
  }
  if (dim != t.dim) {                                        // CHECK_EQ(dim_t, N_)  :85
    mms_set_error("weight_source %s: vectors have %lld dimensions, the layer has num_output = %lld", path, dim, t.dim);
    return MMS_E_INVALID;
  }
  for (long long w = 0; w < vocab; ++w) {
    for (;;) {                                               // the word: up to the next blank (:89-93)
      const int c = fgetc(f);
      if (c == EOF || c == ' ') break;
    }
    if (feof(f)) {
      mms_set_error("weight_source %s: file ends after %lld of %lld words", path, w, vocab);
      return MMS_E_INVALID;
    }
    OVERRUN_CHECK();
    for (long long i = 0; i < dim; ++i) {
      float v = 0.f;
      if (fread(&v, sizeof(float), 1, f) != 1) {
        mms_set_error("weight_source %s: file ends inside the vector of word %lld", path, w);
        return MMS_E_INVALID;
      }
      t.put(v);
    }
  }
  return 0;
}

template <typename T>
int load_weight_source(const char* path, T* table_host, long long rows, long long dim, long long* rows_loaded) {
  MMS_REQUIRE(path && table_host && rows > 0 && dim > 0, MMS_E_INVALID, "bad argument");
  const std::string name(path);
  const bool txt = ends_with(name, "txt"), all = ends_with(name, "all");
  File file(path, (txt || all) ? "r" : "rb");
  if (!file.f) {
    mms_set_error("weight_source %s: %s", path, strerror(errno));
    return MMS_E_INVALID;
  }
  Table<T> t{table_host, rows, dim};
  int rc = txt ? load_glove_txt(file.f, t, path) : all ? load_all(file.f, t, path) : load_word2vec_bin(file.f, t, path);
  if (rows_loaded) *rows_loaded = t.filled / dim;
  return rc;
}

}  // namespace

extern "C" {
int mms_load_weight_source_f32(const char* path, float* table_host, long long input_dim, long long num_output,
                               long long* rows_loaded) {
  return load_weight_source<float>(path, table_host, input_dim, num_output, rows_loaded);
}
int mms_load_weight_source_f64(const char* path, double* table_host, long long input_dim, long long num_output,
                               long long* rows_loaded) {
  return load_weight_source<double>(path, table_host, input_dim, num_output, rows_loaded);
}
}

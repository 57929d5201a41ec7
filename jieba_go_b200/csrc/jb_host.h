// Host-side model of the reference's L0/L1 layers (loaders + tables), producing the flat
// images that are uploaded to HBM.  Plain C++ (no CUDA).
#pragma once
#include <stdint.h>

#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/jieba_b200.h"
#include "jb_common.h"

namespace jb {

// termFreq map[string]int + size (tokenizer.go:381-387) in insertion order
struct HostDict {
  std::vector<std::string> keys;
  std::vector<int64_t> freq;
  std::unordered_map<std::string, uint32_t> index;
  int64_t size = 0;
  // flat view (rebuilt on demand)
  std::vector<uint8_t> blob;
  std::vector<uint32_t> off;
  bool flat_valid = false;

  bool has(const std::string& k) const { return index.find(k) != index.end(); }
  void set(const std::string& k, int64_t v);  // termFreq[k] = v
  void flatten();
};

struct HostEmit {
  std::vector<uint8_t> state;
  std::vector<uint32_t> rune;
  std::vector<double> logp;
};

struct TableImage {
  std::vector<JbFirst> first;
  std::vector<JbEntry> entries;
  std::vector<double> emit;
  std::vector<uint32_t> emit_supp_rune;
  std::vector<double> emit_supp;
  std::vector<uint32_t> han_bits;
  std::vector<uint32_t> supp_lo, supp_hi;
  double neg_log_total = 0;
  double start[4];
  double trans[4][2];
  uint32_t max_delta = 1;
  uint64_t n_han_keys = 0, n_dropped_keys = 0, n_unreachable_keys = 0;
};

double go_log(double x);
bool is_han(uint32_t cp, int unicode_version);
// strict UTF-8 decode (Go rules); returns width, 0 at end; ill-formed -> rune 0xFFFD width 1
int decode_rune(const uint8_t* b, uint64_t i, uint64_t end, uint32_t* r);

int load_dict_text(const uint8_t* data, uint64_t len, int mode, HostDict& d, std::string& err);
int load_dict_gob(const uint8_t* data, uint64_t len, HostDict& d, std::string& err);
int load_emit_json(const uint8_t* data, uint64_t len, HostEmit& e, std::string& err);
int read_file(const char* path, std::vector<uint8_t>& out, std::string& err);
void hmm_defaults(jb_hmm_desc* h);

int build_tables(const jb_dict_desc* dict, const jb_hmm_desc* hmm, int unicode_version, TableImage& img,
                 std::string& err);

// Cached table image: what prefix_dictionary.gob is to dict.txt in the reference (tokenizer.go:439-458, a pre-built
// form of the parsed dictionary), one step further -- the device tables themselves.  `key` = SHA-256 of everything the
// image depends on; a file with another key, another format version or a damaged payload is refused (JB_EFORMAT).
struct Sha256 {
  uint32_t h[8];
  uint64_t len = 0;
  uint8_t buf[64];
  size_t fill = 0;
  Sha256();
  void update(const void* data, size_t n);
  void finish(uint8_t out[32]);
};
int table_image_save(const TableImage& img, const uint8_t key[32], const char* path, std::string& err);
int table_image_load(const char* path, const uint8_t key[32], TableImage& img, std::string& err);

}  // namespace jb

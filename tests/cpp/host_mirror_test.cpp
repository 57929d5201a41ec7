// Driver for include/jieba_b200.hpp: replays a case file written by tests/test_cpp_host_mirror.py (expected tokens come
// from the oracle) through the C++ mirror of the reference's interface -- Cut / CutParallel / CutBatch / AddWord -- the
// way the reference's own tests call them (tokenizer_test.go TestCut, TestCutParallel, TestAddWord).
//
//   host_mirror_test <case file>     exit 0 = every case identical, 1 = mismatch, 2 = bad case file
//   host_mirror_test --no-device     exit 0 iff construction fails loudly with JB_ECUDA (no CPU fallback)
//
// Case file: one record per line, fields separated by one space, strings hex-encoded ("-" = empty):
//   dict <path> <mode>          emit <path>          unicode <13|15>
//   cut <hmm> <text> <tok,tok,...>
//   par <hmm> <workers> <ordered> <text> <tok,tok,...>
//   add <word> <freq>           freq <word> <expected count>
//   batch                       (every cut record so far again, through CutBatch, both hmm settings apart)
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>

#include "jieba_b200.hpp"

using jieba_b200::Tokenizer;

static std::string unhex(const std::string& h) {
  if (h == "-") return "";
  std::string out;
  for (size_t i = 0; i + 1 < h.size(); i += 2) out.push_back((char)std::stoi(h.substr(i, 2), nullptr, 16));
  return out;
}
static std::vector<std::string> unhex_list(const std::string& f) {
  std::vector<std::string> out;
  if (f == "-") return out;
  std::stringstream ss(f);
  std::string item;
  while (std::getline(ss, item, ',')) out.push_back(unhex(item));
  return out;
}
static std::string show(const std::vector<std::string>& v) {
  std::string s;
  for (auto& t : v) s += "[" + t + "]";
  return s;
}

struct CutCase {
  bool hmm;
  std::string text;
  std::vector<std::string> want;
};

int main(int argc, char** argv) {
  if (argc == 2 && std::string(argv[1]) == "--no-device") {
    try {
      auto tk = Tokenizer::FromMemory("\xE7\x94\xB2 100\n", JB_DICT_PREFIX_MODE, "");
    } catch (const jieba_b200::Error& e) {
      std::printf("construction failed as it must: %s\n", e.what());
      return e.code == JB_ECUDA ? 0 : 1;
    }
    std::printf("construction succeeded: a device is present\n");
    return 3;
  }
  if (argc != 2) return 2;
  std::ifstream in(argv[1]);
  if (!in) return 2;
  std::string dict_path, emit_path, line;
  int mode = JB_DICT_PREFIX_MODE, uni = 15;
  std::unique_ptr<Tokenizer> tk;
  std::vector<CutCase> seen;
  int checked = 0, bad = 0;
  auto need = [&]() {
    if (tk) return;
    jieba_b200::Options o;
    o.unicode_version = uni;
    o.emit_json_path = emit_path;
    o.dict_mode = mode;
    tk = Tokenizer::NewTokenizer(dict_path, o);
  };
  auto compare = [&](const char* what, const std::string& text, const std::vector<std::string>& got, const std::vector<std::string>& want) {
    checked++;
    if (got == want) return;
    bad++;
    if (bad <= 10) std::printf("MISMATCH %s text=%s\n  got  %s\n  want %s\n", what, text.c_str(), show(got).c_str(), show(want).c_str());
  };
  try {
    while (std::getline(in, line)) {
      if (line.empty() || line[0] == '#') continue;
      std::stringstream ss(line);
      std::string op;
      ss >> op;
      if (op == "dict") {
        ss >> dict_path >> mode;
      } else if (op == "emit") {
        ss >> emit_path;
      } else if (op == "unicode") {
        ss >> uni;
      } else if (op == "cut") {
        int hmm;
        std::string t, w;
        ss >> hmm >> t >> w;
        need();
        CutCase c{hmm != 0, unhex(t), unhex_list(w)};
        compare("Cut", c.text, tk->Cut(c.text, c.hmm), c.want);
        seen.push_back(c);
      } else if (op == "par") {
        int hmm, workers, ordered;
        std::string t, w;
        ss >> hmm >> workers >> ordered >> t >> w;
        need();
        const std::string text = unhex(t);
        compare("CutParallel", text, tk->CutParallel(text, hmm != 0, workers, ordered != 0), unhex_list(w));
      } else if (op == "add") {
        std::string w;
        int freq;
        ss >> w >> freq;
        need();
        tk->AddWord(unhex(w), freq);
      } else if (op == "freq") {
        std::string w;
        long long want;
        ss >> w >> want;
        need();
        int64_t got = -1;
        const bool found = tk->Lookup(unhex(w), &got);
        checked++;
        if (!found || got != want) {
          bad++;
          std::printf("MISMATCH freq of %s: got %lld (found %d) want %lld\n", unhex(w).c_str(), (long long)got, (int)found, want);
        }
      } else if (op == "batch") {
        need();
        for (int hmm = 0; hmm < 2; hmm++) {
          std::vector<std::string_view> texts;
          std::vector<const CutCase*> which;
          for (auto& c : seen)
            if (c.hmm == (hmm != 0)) {
              texts.push_back(c.text);
              which.push_back(&c);
            }
          auto got = tk->CutBatch(texts, hmm != 0);
          for (size_t i = 0; i < which.size(); i++) compare("CutBatch", which[i]->text, got[i], which[i]->want);
        }
      } else {
        std::printf("bad record: %s\n", line.c_str());
        return 2;
      }
    }
  } catch (const std::exception& e) {
    std::printf("EXCEPTION %s\n", e.what());
    return 1;
  }
  std::printf("%d checks, %d mismatches\n", checked, bad);
  return bad ? 1 : 0;
}

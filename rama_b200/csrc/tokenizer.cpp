// tokenizer.cpp — host side of the text boundary of the decode path: the llama2.c tokenizer.bin reader and
// the greedy BPE encode / piece decode of the reference (engine/src/tokenizer/bpe.rs), kept byte for byte
// so that `generate(prompt text) -> text` (mod.rs:169-206) has the same ids in and the same bytes out.
// Pure host code (the reference tokenizes on the CPU too); nothing here touches the GPU.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/rama_b200.h"

struct rama_tokenizer {
  std::vector<std::string> vocab;                    // bpe.rs:11
  std::vector<float> scores;                         // bpe.rs:12
  std::unordered_map<std::string, int32_t> ids;      // bpe.rs:13 (insert overwrites: the last duplicate wins)
  uint32_t max_token_length = 0;                     // bpe.rs:15
};

extern "C" int rama_set_error(int code, const char* msg);  // api.cu (thread-local message)

// ≙ Tokenizer::new(path, vocab_size) (bpe.rs:19-45): u32 max_token_length, then per token f32 score,
// i32 byte length, bytes.  The reference unwraps every read and String::from_utf8: malformed files are errors.
extern "C" int rama_tokenizer_load(const char* path, int32_t vocab_size, rama_tokenizer** out) {
  if (!path || !out || vocab_size <= 0) return rama_set_error(RAMA_E_INVALID, "tokenizer: bad argument");
  FILE* f = fopen(path, "rb");
  if (!f) return rama_set_error(RAMA_E_IO, "tokenizer: cannot open file");
  rama_tokenizer* t = new rama_tokenizer();
  bool ok = fread(&t->max_token_length, 4, 1, f) == 1;
  for (int32_t i = 0; ok && i < vocab_size; ++i) {
    float score;
    int32_t len;
    ok = fread(&score, 4, 1, f) == 1 && fread(&len, 4, 1, f) == 1 && len >= 0 && len < (1 << 20);
    if (!ok) break;
    std::string s((size_t)len, '\0');
    ok = len == 0 || fread(&s[0], 1, (size_t)len, f) == (size_t)len;
    if (!ok) break;
    t->scores.push_back(score);
    t->ids[s] = i;
    t->vocab.push_back(std::move(s));
  }
  fclose(f);
  if (!ok) {
    delete t;
    return rama_set_error(RAMA_E_IO, "tokenizer: file shorter than vocab_size entries");
  }
  *out = t;
  return RAMA_OK;
}

extern "C" int rama_tokenizer_free(rama_tokenizer* t) {
  delete t;
  return RAMA_OK;
}

extern "C" int rama_tokenizer_info(const rama_tokenizer* t, int32_t* vocab_size, int32_t* max_token_length) {
  if (!t) return rama_set_error(RAMA_E_INVALID, "tokenizer: NULL");
  if (vocab_size) *vocab_size = (int32_t)t->vocab.size();
  if (max_token_length) *max_token_length = (int32_t)t->max_token_length;
  return RAMA_OK;
}

// one UTF-8 scalar starting at s[i]; returns its byte length (malformed bytes are passed through one by one:
// a Rust &str is always valid UTF-8, so this only matters for callers that hand over other encodings)
static size_t utf8_len(const std::string& s, size_t i) {
  const unsigned char c = (unsigned char)s[i];
  size_t n = c < 0x80 ? 1 : (c >> 5) == 0x6 ? 2 : (c >> 4) == 0xE ? 3 : (c >> 3) == 0x1E ? 4 : 1;
  if (i + n > s.size()) n = 1;
  for (size_t k = 1; k < n; ++k)
    if (((unsigned char)s[i + k] & 0xC0) != 0x80) return 1;
  return n;
}
static uint32_t utf8_cp(const std::string& s, size_t i, size_t n) {
  const unsigned char* p = (const unsigned char*)s.data() + i;
  switch (n) {
    case 1: return p[0];
    case 2: return ((p[0] & 0x1Fu) << 6) | (p[1] & 0x3Fu);
    case 3: return ((p[0] & 0x0Fu) << 12) | ((p[1] & 0x3Fu) << 6) | (p[2] & 0x3Fu);
    default: return ((p[0] & 0x07u) << 18) | ((p[1] & 0x3Fu) << 12) | ((p[2] & 0x3Fu) << 6) | (p[3] & 0x3Fu);
  }
}
// char::is_whitespace (the Unicode White_Space property), what str::trim strips
static bool is_ws(uint32_t c) {
  return (c >= 0x09 && c <= 0x0D) || c == 0x20 || c == 0x85 || c == 0xA0 || c == 0x1680 || (c >= 0x2000 && c <= 0x200A) ||
         c == 0x2028 || c == 0x2029 || c == 0x202F || c == 0x205F || c == 0x3000;
}

// ≙ Tokenizer::encode (bpe.rs:50-97): trim, drop '\n', one token per char (a char missing from the vocabulary
// panics there: RAMA_E_INVALID here), then repeatedly merge the adjacent pair whose concatenation is the
// highest-scoring vocabulary entry (strict >, so the leftmost wins ties) until no pair merges.
// An input that leaves no token panics in the reference (usize underflow at bpe.rs:66): RAMA_E_INVALID.
extern "C" int rama_tokenizer_encode(const rama_tokenizer* t, const char* text, int32_t* out, int32_t cap, int32_t* n_out) {
  if (!t || !text || !n_out) return rama_set_error(RAMA_E_INVALID, "tokenizer: NULL argument");
  const std::string s(text);
  // trim
  std::vector<std::pair<size_t, size_t>> chars;  // (offset, byte length)
  for (size_t i = 0; i < s.size();) {
    const size_t n = utf8_len(s, i);
    chars.push_back({i, n});
    i += n;
  }
  size_t lo = 0, hi = chars.size();
  while (lo < hi && is_ws(utf8_cp(s, chars[lo].first, chars[lo].second))) ++lo;
  while (hi > lo && is_ws(utf8_cp(s, chars[hi - 1].first, chars[hi - 1].second))) --hi;
  std::vector<int32_t> tokens;
  for (size_t k = lo; k < hi; ++k) {
    if (chars[k].second == 1 && s[chars[k].first] == '\n') continue;  // bpe.rs:54
    auto it = t->ids.find(s.substr(chars[k].first, chars[k].second));
    if (it == t->ids.end()) return rama_set_error(RAMA_E_INVALID, "tokenizer: character not in the vocabulary (the reference panics, bpe.rs:55)");
    tokens.push_back(it->second);
  }
  if (tokens.empty()) return rama_set_error(RAMA_E_INVALID, "tokenizer: nothing to encode (the reference underflows and panics, bpe.rs:66)");
  std::string buf;
  for (;;) {
    float best_score = -1e10f;
    int32_t best_id = -1;
    size_t best_idx = (size_t)-1;
    for (size_t i = 0; i + 1 < tokens.size(); ++i) {
      buf.assign(t->vocab[tokens[i]]);
      buf.append(t->vocab[tokens[i + 1]]);
      auto it = t->ids.find(buf);
      if (it != t->ids.end() && t->scores[it->second] > best_score) {
        best_score = t->scores[it->second];
        best_id = it->second;
        best_idx = i;
      }
    }
    if (best_idx == (size_t)-1) break;
    tokens[best_idx] = best_id;
    tokens.erase(tokens.begin() + best_idx + 1);
  }
  *n_out = (int32_t)tokens.size();
  if (out) {
    if ((int32_t)tokens.size() > cap) return rama_set_error(RAMA_E_INVALID, "tokenizer: output buffer too small");
    memcpy(out, tokens.data(), tokens.size() * sizeof(int32_t));
  }
  return RAMA_OK;
}

static int hex_val(char c) {
  if (c >= '0' && c <= '9') return c - '0';
  if (c >= 'a' && c <= 'f') return c - 'a' + 10;
  if (c >= 'A' && c <= 'F') return c - 'A' + 10;
  return -1;
}

// ≙ decode(tokenizer.vocab[token]) (bpe.rs:102-116): a piece containing "<s>" prints nothing; a piece that
// starts with '<' and ends with '>' is taken as "<0xAB>" and becomes char::from(0xAB) (Latin-1 → UTF-8);
// such a piece without two hex digits at bytes 3..5 panics in the reference ("</s>", "<unk>"): RAMA_E_INVALID.
extern "C" int rama_tokenizer_decode(const rama_tokenizer* t, int32_t token, char* out, int32_t cap, int32_t* n_out) {
  if (!t || !n_out) return rama_set_error(RAMA_E_INVALID, "tokenizer: NULL argument");
  if (token < 0 || (size_t)token >= t->vocab.size()) return rama_set_error(RAMA_E_INVALID, "tokenizer: token outside the vocabulary");
  const std::string& p = t->vocab[(size_t)token];
  std::string r;
  if (p.find("<s>") != std::string::npos) {
    r.clear();
  } else if (!p.empty() && p.front() == '<' && p.back() == '>') {
    int h = -1, l = -1;
    if (p.size() >= 5) {
      // u8::from_str_radix accepts an optional leading '+'
      if (p[3] == '+') { h = 0; l = hex_val(p[4]); }
      else { h = hex_val(p[3]); l = hex_val(p[4]); }
    }
    if (h < 0 || l < 0) return rama_set_error(RAMA_E_INVALID, "tokenizer: piece looks like <0x..> but has no hex byte (the reference panics, bpe.rs:110)");
    const unsigned c = (unsigned)(h * 16 + l);
    if (c < 0x80) r.push_back((char)c);
    else { r.push_back((char)(0xC0 | (c >> 6))); r.push_back((char)(0x80 | (c & 0x3F))); }
  } else {
    r = p;
  }
  *n_out = (int32_t)r.size();
  if (out) {
    if ((int32_t)r.size() > cap) return rama_set_error(RAMA_E_INVALID, "tokenizer: output buffer too small");
    memcpy(out, r.data(), r.size());
  }
  return RAMA_OK;
}

// refset.cu — host-side loading of the V/J references and motif tables, and the device tables built from them.
//
// Replaces what `bwa index` + `FASTA.Reader` give the reference (catt.jl:198-205; resource/**.pac/.ann/.amb):
//   * the 2-strand text T = cat + revcomp(cat) that bwa's FM-index searches, 2-bit packed;
//   * a 10-mer presence bitmap + CSR position index over T (the seed table the seeding kernel probes);
//   * per-record offsets/lengths and the FASTA bases with N preserved (what real_score, catt.jl:69-84, sees).
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <omp.h>
#include <zlib.h>

#include <algorithm>

#include "gz.h"
#include "internal.cuh"

namespace catt {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  set_error("CUDA error %s (%d) at %s:%d in %s", cudaGetErrorString(e), (int)e, file, line, what);
  return CATT_E_CUDA;
}

// lrand48 as bwa's index builder calls it (srand48(11), one draw per reference N) when no .pac is at hand.
struct Lrand48 {
  uint64_t x = ((uint64_t)11 << 16) | 0x330E;
  uint32_t next() { x = (x * 0x5DEECE66DULL + 0xB) & ((1ULL << 48) - 1); return (uint32_t)(x >> 17); }
};

int RefSet::load(const std::string& fasta) {
  std::string txt;
  if (!slurp_file(fasta, txt)) { set_error("cannot read reference FASTA %s", fasta.c_str()); return CATT_E_IO; }
  names.clear(); off.clear(); len.clear(); fa.clear(); T.clear(); L = 0; maxlen = 0;
  size_t p = 0;
  int cur = -1;
  while (p < txt.size()) {
    size_t e = txt.find('\n', p);
    if (e == std::string::npos) e = txt.size();
    size_t b = p, q = e;
    p = e + 1;
    while (q > b && (txt[q - 1] == '\r' || txt[q - 1] == ' ' || txt[q - 1] == '\t')) q--;
    if (q == b) continue;
    if (txt[b] == '>') {
      size_t k = b + 1;
      while (k < q && txt[k] != ' ' && txt[k] != '\t') k++;
      names.emplace_back(txt, b + 1, k - b - 1);
      off.push_back(L);
      len.push_back(0);
      cur = (int)names.size() - 1;
    } else if (cur >= 0) {
      for (size_t k = b; k < q; k++) fa.push_back(nt_code(txt[k]));
      len[cur] += (int)(q - b);
      L += (int)(q - b);
    }
  }
  if (names.empty() || L == 0) { set_error("no records in reference FASTA %s", fasta.c_str()); return CATT_E_IO; }
  if ((int)names.size() > MAX_REC) { set_error("%s: %zu records exceed the limit of %d", fasta.c_str(), names.size(), MAX_REC); return CATT_E_LIMIT; }
  for (int l : len) maxlen = std::max(maxlen, l);
  if (maxlen > MAX_REC_LEN) { set_error("%s: record of %d nt exceeds the limit of %d", fasta.c_str(), maxlen, MAX_REC_LEN); return CATT_E_LIMIT; }
  // alignment bases: N -> what the bwa index holds
  std::string pac;
  bool have_pac = slurp_file(fasta + ".pac", pac) && (int)pac.size() >= (L + 3) / 4;
  Lrand48 rng;
  T.resize(2 * (size_t)L);
  for (int i = 0; i < L; i++) {
    uint8_t c = fa[i];
    if (c > 3) {
      uint8_t r = (uint8_t)(rng.next() & 3);
      c = have_pac ? (uint8_t)((pac[i >> 2] >> ((~i & 3) << 1)) & 3) : r;
    }
    T[i] = c;
    T[2 * (size_t)L - 1 - i] = (uint8_t)(3 - c);
  }
  return CATT_OK;
}

template <class V>
static int to_device(const V& v, const void** out, std::vector<void*>& allocs) {
  void* d = nullptr;
  size_t bytes = v.size() * sizeof(typename V::value_type);
  CATT_CUDA(cudaMalloc(&d, std::max<size_t>(bytes, 16)));
  allocs.push_back(d);
  CATT_CUDA(cudaMemcpy(d, v.data(), bytes, cudaMemcpyHostToDevice));
  *out = d;
  return CATT_OK;
}

int RefSet::upload() {
  const int n2 = 2 * L;
  // 2-bit text, padded by 2 words so 64-bit window reads stay in bounds
  std::vector<uint32_t> t2(t2_words(L), 0);
  for (int i = 0; i < n2; i++) t2[i >> 4] |= (uint32_t)T[i] << ((i & 15) * 2);
  // 10-mer index; key = sum code[p+j] << 2j (LSB-first, like the packed reads)
  const int NK = 1 << 20;
  std::vector<uint32_t> kstart(NK + 1, 0), bitmap(NK / 32, 0);
  auto key_at = [&](int p) { uint32_t k = 0; for (int j = 0; j < SEED_K; j++) k |= (uint32_t)T[p + j] << (2 * j); return k; };
  for (int p = 0; p + SEED_K <= n2; p++) kstart[key_at(p) + 1]++;
  for (int i = 0; i < NK; i++) kstart[i + 1] += kstart[i];
  std::vector<uint32_t> kpos(std::max(1, n2 - SEED_K + 1));
  // every key's occurrence list is grouped by the text base BEFORE the occurrence (groups 0..3, then "none" for p = 0):
  // a window whose previous read base is b starts a new maximal run exactly at the occurrences outside group b, so the
  // seeding kernel skips group b wholesale.  ksub[key] = cumulative sizes of groups 0..3 (4 x u16).
  std::vector<unsigned long long> ksub(NK, 0ull);
  {
    std::vector<uint32_t> gcount((size_t)NK * 4, 0);
    auto group_of = [&](int p) { return p == 0 ? 4 : (int)T[p - 1]; };
    for (int p = 0; p + SEED_K <= n2; p++) { const int g = group_of(p); if (g < 4) gcount[(size_t)key_at(p) * 4 + g]++; }
    std::vector<uint32_t> fill((size_t)NK * 5);
    for (int k = 0; k < NK; k++) {
      if (kstart[k + 1] == kstart[k]) continue;
      uint32_t o = kstart[k], cum = 0;
      unsigned long long sub = 0;
      for (int g = 0; g < 4; g++) {
        fill[(size_t)k * 5 + g] = o + cum;
        cum += gcount[(size_t)k * 4 + g];
        if (cum > 0xFFFFu) { set_error("10-mer with more than 65535 occurrences in the reference"); return CATT_E_LIMIT; }
        sub |= (unsigned long long)cum << (16 * g);
      }
      fill[(size_t)k * 5 + 4] = o + cum;
      ksub[k] = sub;
    }
    for (int p = 0; p + SEED_K <= n2; p++) {
      const uint32_t k = key_at(p);
      kpos[fill[(size_t)k * 5 + group_of(p)]++] = (uint32_t)p;
      bitmap[k >> 5] |= 1u << (k & 31);
    }
  }
  std::vector<uint16_t> pos2rid(L);
  for (size_t r = 0; r < names.size(); r++) for (int i = 0; i < len[r]; i++) pos2rid[off[r] + i] = (uint16_t)r;
  // byte-identical records share one DP: canon[rid*2+strand] = (first identical rid)*2+strand
  std::vector<uint16_t> canon(names.size() * 2);
  for (size_t r = 0; r < names.size(); r++) {
    size_t f = r;
    for (size_t q = 0; q < r; q++)
      if (len[q] == len[r] && std::equal(T.begin() + off[q], T.begin() + off[q] + len[q], T.begin() + off[r])) { f = q; break; }
    canon[2 * r] = (uint16_t)(2 * f); canon[2 * r + 1] = (uint16_t)(2 * f + 1);
  }
  std::vector<uint8_t> tb(T.begin(), T.begin() + L);
  tb.resize(L + 64, 0);
  int rc;
  if ((rc = to_device(t2, (const void**)&dev.T2, allocs))) return rc;
  if ((rc = to_device(tb, (const void**)&dev.Tb, allocs))) return rc;
  if ((rc = to_device(fa, (const void**)&dev.fa, allocs))) return rc;
  if ((rc = to_device(bitmap, (const void**)&dev.bitmap, allocs))) return rc;
  if ((rc = to_device(kstart, (const void**)&dev.kstart, allocs))) return rc;
  if ((rc = to_device(kpos, (const void**)&dev.kpos, allocs))) return rc;
  if ((rc = to_device(ksub, (const void**)&dev.ksub, allocs))) return rc;
  if ((rc = to_device(pos2rid, (const void**)&dev.pos2rid, allocs))) return rc;
  if ((rc = to_device(canon, (const void**)&dev.canon, allocs))) return rc;
  if ((rc = to_device(off, (const void**)&dev.rec_off, allocs))) return rc;
  if ((rc = to_device(len, (const void**)&dev.rec_len, allocs))) return rc;
  dev.L = L; dev.nrec = (int)names.size(); dev.maxlen = maxlen; dev.CW = ((int)names.size() * 2 + 31) / 32;
  return CATT_OK;
}

void RefSet::release() {
  for (void* p : allocs) cudaFree(p);
  allocs.clear();
}

// ---------------------------------------------------------------------------------------------------------------
// motif regex -> fixed-length alternatives of residue masks.  Every motif in reference.jl:13-116 is a union of
// fixed-length residue patterns (literals, [classes], {n}, (a|b) groups); anything else is rejected loudly.
// ---------------------------------------------------------------------------------------------------------------
namespace {
typedef std::vector<std::vector<uint32_t>> Alts;
inline int aa_bit(char c) { return (c >= 'A' && c <= 'Z') ? c - 'A' : (c == '*' ? 26 : -1); }
struct Rx {
  const char* s; size_t n, p = 0; bool ok = true;
  static Alts cat(const Alts& a, const Alts& b) {
    Alts r;
    for (auto& x : a) for (auto& y : b) { r.push_back(x); r.back().insert(r.back().end(), y.begin(), y.end()); }
    return r;
  }
  Alts alt() {
    Alts r = seq();
    while (p < n && s[p] == '|') { p++; Alts t = seq(); r.insert(r.end(), t.begin(), t.end()); }
    return r;
  }
  Alts seq() {
    Alts r{{}};
    while (p < n && s[p] != '|' && s[p] != ')') {
      Alts a = atom();
      if (p < n && s[p] == '{') {
        size_t e = p; while (e < n && s[e] != '}') e++;
        if (e >= n) { ok = false; return r; }
        int cnt = atoi(std::string(s + p + 1, e - p - 1).c_str());
        p = e + 1;
        Alts rep{{}};
        for (int i = 0; i < cnt; i++) rep = cat(rep, a);
        a = rep;
      } else if (p < n && (s[p] == '*' || s[p] == '+' || s[p] == '?')) { ok = false; return r; }
      r = cat(r, a);
      if (r.size() > 4096) { ok = false; return r; }
    }
    return r;
  }
  Alts atom() {
    char c = s[p];
    if (c == '(') { p++; Alts r = alt(); if (p < n && s[p] == ')') p++; else ok = false; return r; }
    if (c == '[') {
      p++;
      uint32_t m = 0;
      while (p < n && s[p] != ']') {
        if (p + 2 < n && s[p + 1] == '-' && s[p + 2] != ']') {
          for (int x = s[p]; x <= s[p + 2]; x++) { int b = aa_bit((char)x); if (b >= 0) m |= 1u << b; }
          p += 3;
        } else { int b = aa_bit(s[p]); if (b >= 0) m |= 1u << b; p++; }
      }
      if (p >= n) ok = false; else p++;
      return Alts{{m}};
    }
    if (c == '.') { p++; return Alts{{0x7FFFFFFu}}; }
    p++;
    int b = aa_bit(c);
    return Alts{{b >= 0 ? 1u << b : 0u}};   // a character outside the residue alphabet ("place_holder") never matches
  }
};
}  // namespace

int compile_motif(const char* re, MotifDev* out) {
  memset(out, 0, sizeof *out);
  Rx rx{re, strlen(re)};
  Alts a = rx.alt();
  if (!rx.ok || rx.p != rx.n || a.empty()) { set_error("motif '%s': unsupported regex syntax", re); return CATT_E_ARG; }
  // drop alternatives that can never match, then merge duplicates
  Alts live;
  size_t L0 = a[0].size();
  for (auto& x : a) {
    if (x.size() != L0) { set_error("motif '%s' is not fixed-length", re); return CATT_E_ARG; }
    bool dead = false;
    for (uint32_t m : x) dead |= (m == 0);
    if (!dead && std::find(live.begin(), live.end(), x) == live.end()) live.push_back(x);
  }
  if (live.empty()) { out->len = (int)std::min<size_t>(L0, MOTIF_MAX_LEN); out->nalt = 0; return CATT_OK; }
  if (L0 > MOTIF_MAX_LEN || L0 == 0) { set_error("motif '%s': length %zu unsupported", re, L0); return CATT_E_ARG; }
  if (live.size() > MOTIF_MAX_ALT) { set_error("motif '%s': %zu alternatives exceed %d", re, live.size(), MOTIF_MAX_ALT); return CATT_E_ARG; }
  out->len = (int)L0; out->nalt = (int)live.size();
  for (size_t i = 0; i < live.size(); i++)
    for (size_t k = 0; k < L0; k++)
      for (int r = 0; r < 27; r++)
        if ((live[i][k] >> r) & 1u) out->alt[k][r] |= 1ull << i;
  return CATT_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// read packing (host): ASCII -> 2-bit words + N mask, straight into pinned staging (api.cu pipelines it chunk by
// chunk under the kernels of the previous chunk).
// ---------------------------------------------------------------------------------------------------------------
namespace {
struct CodeLut {
  uint8_t v[256];
  CodeLut() { for (int i = 0; i < 256; i++) v[i] = nt_code((char)i); }
};
const CodeLut LUT;
}  // namespace

int pack_layout(int64_t n, const int64_t* seq_off, uint32_t* off, int32_t* len, int* max_len) {
  const int nt = std::max(1, omp_get_max_threads());
  std::vector<uint64_t> part((size_t)nt + 1, 0);
  std::vector<int> pmax((size_t)nt, 0);
  int64_t bad = -1;
  #pragma omp parallel num_threads(nt)
  {
    const int t = omp_get_thread_num();
    const int64_t lo = n * t / nt, hi = n * (t + 1) / nt;
    uint64_t w = 0; int mx = 0;
    for (int64_t i = lo; i < hi; i++) {
      const int64_t l = seq_off[i + 1] - seq_off[i];
      if (l > MAX_READ || l < 0) {
        #pragma omp critical
        { if (bad < 0 || i < bad) bad = i; }
        continue;
      }
      len[i] = (int32_t)l; mx = std::max(mx, (int)l); w += read_words((int)l);
    }
    part[t + 1] = w; pmax[t] = mx;
    #pragma omp barrier
    #pragma omp single
    { for (int k = 0; k < nt; k++) part[k + 1] += part[k]; }
    if (bad < 0 && part[nt] <= 0xFFFFFFF0ull) {
      uint64_t o = part[t];
      for (int64_t i = lo; i < hi; i++) { off[i] = (uint32_t)o; o += read_words(len[i]); }
    }
  }
  if (bad >= 0) {
    set_error("read %lld is %lld nt long; the limit is %d", (long long)bad, (long long)(seq_off[bad + 1] - seq_off[bad]), MAX_READ);
    return CATT_E_LIMIT;
  }
  if (part[nt] > 0xFFFFFFF0ull) { set_error("batch too large: packed reads exceed 2^32 words"); return CATT_E_LIMIT; }
  off[n] = (uint32_t)part[nt];
  int mx = 0;
  for (int v : pmax) mx = std::max(mx, v);
  *max_len = mx;
  return CATT_OK;
}

void pack_range(int64_t i0, int64_t i1, const char* seqs, const int64_t* seq_off, const uint32_t* off, uint32_t* dst) {
  const uint32_t base = i0 < i1 ? off[i0] : 0;
  #pragma omp parallel for schedule(static)
  for (int64_t i = i0; i < i1; i++) {
    const unsigned char* s = reinterpret_cast<const unsigned char*>(seqs + seq_off[i]);
    const int l = (int)(seq_off[i + 1] - seq_off[i]);
    uint32_t* b = dst + (off[i] - base);
    uint32_t* m = b + read_base_words(l);
    for (int k = 0; k < l; k += 32) {
      const int cnt = std::min(32, l - k);
      uint64_t bw = 0; uint32_t mw = 0;
      for (int j = 0; j < cnt; j++) {
        const uint32_t c = LUT.v[s[k + j]];
        bw |= (uint64_t)(c & 3u) << (2 * j);
        mw |= (c >> 2) << j;
      }
      b[k >> 4] = (uint32_t)bw;
      if (k + 16 < l) b[(k >> 4) + 1] = (uint32_t)(bw >> 32);
      m[k >> 5] = mw;
    }
  }
}

int pack_reads(int64_t n, const char* seqs, const int64_t* seq_off, PackedReads* out) {
  out->off.resize(n + 1);
  out->len.resize(std::max<int64_t>(n, 1));
  int rc = pack_layout(n, seq_off, out->off.data(), out->len.data(), &out->max_len);
  if (rc) return rc;
  out->len.resize(n);
  out->words.assign((size_t)out->off[n] + 4, 0);
  pack_range(0, n, seqs, seq_off, out->off.data(), out->words.data());
  return CATT_OK;
}

}  // namespace catt

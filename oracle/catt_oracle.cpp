// catt_oracle.cpp — CPU ORACLE for the CATT per-read CDR3 extraction hot path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing in the product path (catt_b200/, libcatt_b200.so) may
// include, link or call this file; only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs use it, as the checker or as the timed CPU baseline.
//
// What it restates (all file:line relative to /root/reference):
//   * map2align                 catt.jl:37-56   (bwa mem 0.7.17 -k10 -A1 -B2 -L0 -T12 | samtools sort -t AS |
//                                                view -F 2308 | tac | awk '!a[$1]++')
//   * real_score/assignV/J      catt.jl:58-132
//   * get_kmer/err_fromsam      catt.jl:153-184
//   * read_alignrs              catt.jl:186-339
//   * ratio/extend_*!/pool      catt.jl:341-497, 544-577
//   * cigar helpers             Jtool.jl:76-124
//   * Extract_Motif2/finder!    Jtool.jl:234-316, 355-389
//   * selBestMatchD             Jtool.jl:21-28, 40-52
// bwa-mem / samtools / Kmers.jl / BioAlignments are third-party and absent from /root/reference
// (pinned only by name: bwa-0.7.17 in resource/TR/hs/TRA/comm.sh:3); their published algorithms are
// restated following SURVEY.md Appendix A.  Parity is PINNED end-to-end by the reference's own
// golden output aha.TRB.CDR3.CATT.csv for testSample.fq (tests/test_oracle_golden.py) and by its
// Dregion column for selBestMatchD.  Per-read AS/POS/CIGAR have no shipped golden: "parity unpinned"
// at that granularity.
//
// Build: see oracle/Makefile (g++ -O2 -fopenmp -shared).  C interface at the bottom of the file.
#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>
#include <zlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// ----------------------------------------------------------------------------------------------
// small helpers
// ----------------------------------------------------------------------------------------------
inline uint8_t nt_code(char c) {
  switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return 4;
  }
}
const char NT[6] = "ACGTN";

// bwa's 64-bit mix used to break ties between equal-score hits.
inline uint64_t hash_64(uint64_t key) {
  key += ~(key << 32);
  key ^= (key >> 22);
  key += ~(key << 13);
  key ^= (key >> 8);
  key += (key << 3);
  key ^= (key >> 15);
  key += ~(key << 27);
  key ^= (key >> 31);
  return key;
}

// glibc/BSD lrand48 as bwa index uses it (srand48(11); N -> lrand48()&3) to fill reference N's.
struct Rand48 {
  uint64_t x;
  explicit Rand48(uint32_t seed) { x = ((uint64_t)seed << 16) | 0x330E; }
  uint32_t next() {
    x = (x * 0x5DEECE66DULL + 0xB) & ((1ULL << 48) - 1);
    return (uint32_t)(x >> 17);
  }
};

// ----------------------------------------------------------------------------------------------
// reference set (one FASTA = the V or the J "genome" bwa indexes)          SURVEY A2
// ----------------------------------------------------------------------------------------------
struct RefSet {
  std::vector<std::string> names;
  std::vector<int> off, len;
  int L = 0;                  // forward length
  std::vector<uint8_t> T;     // 2L alignment bases 0..3: forward + reverse complement
  std::vector<uint8_t> fa;    // L FASTA bases 0..4 (N kept; the Julia side sees these)
  std::vector<int> pos2rid;   // L
  std::vector<int> kstart;    // 10-mer CSR index over T (windows may cross record / strand boundaries,
  std::vector<int> kpos;      // exactly like the FM-index of the concatenated text)
  int n_from_pac = 0;         // how many N's were resolved from the .pac
  int pac_mismatch = 0;       // non-N bases that disagree with the .pac (should be 0)
};

bool read_file(const std::string& path, std::string& out) {
  gzFile f = gzopen(path.c_str(), "rb");
  if (!f) return false;
  out.clear();
  char buf[1 << 16];
  int n;
  while ((n = gzread(f, buf, sizeof buf)) > 0) out.append(buf, n);
  gzclose(f);
  return true;
}

RefSet* load_refset(const std::string& fasta, int use_pac) {
  std::string txt;
  if (!read_file(fasta, txt)) return nullptr;
  RefSet* rs = new RefSet();
  size_t p = 0;
  std::string seq;
  auto flush = [&]() {
    if (rs->names.empty()) return;
    rs->off.push_back(rs->L);
    rs->len.push_back((int)seq.size());
    for (char c : seq) rs->fa.push_back(nt_code(c));
    rs->L += (int)seq.size();
    seq.clear();
  };
  while (p < txt.size()) {
    size_t e = txt.find('\n', p);
    if (e == std::string::npos) e = txt.size();
    std::string line = txt.substr(p, e - p);
    p = e + 1;
    while (!line.empty() && (line.back() == '\r' || line.back() == ' ')) line.pop_back();
    if (line.empty()) continue;
    if (line[0] == '>') {
      flush();
      size_t k = 1;
      while (k < line.size() && !isspace((unsigned char)line[k])) k++;
      rs->names.push_back(line.substr(1, k - 1));   // FASTA.identifier (catt.jl:202) == bwa RNAME
    } else {
      seq += line;
    }
  }
  flush();
  const int L = rs->L;
  std::vector<uint8_t> fwd(rs->fa);
  // N handling: bwa index writes lrand48()&3 into the .pac (srand48(11)); prefer the shipped .pac.
  std::string pac;
  bool have_pac = use_pac && read_file(fasta + ".pac", pac) && (int)pac.size() >= (L + 3) / 4;
  Rand48 rng(11);
  for (int i = 0; i < L; i++) {
    uint8_t pb = have_pac ? (uint8_t)((pac[i >> 2] >> ((~i & 3) << 1)) & 3) : 0;
    if (fwd[i] > 3) {
      uint8_t r = (uint8_t)(rng.next() & 3);
      fwd[i] = have_pac ? pb : r;
      if (have_pac) rs->n_from_pac++;
      if (have_pac && pb != r) rs->pac_mismatch += 1000;   // lrand48 emulation disagrees with the shipped .pac
    } else if (have_pac && pb != fwd[i]) {
      rs->pac_mismatch++;
    }
  }
  rs->T.resize(2 * (size_t)L);
  for (int i = 0; i < L; i++) {
    rs->T[i] = fwd[i];
    rs->T[2 * L - 1 - i] = (uint8_t)(3 - fwd[i]);
  }
  rs->pos2rid.resize(L);
  for (size_t r = 0; r < rs->names.size(); r++)
    for (int i = 0; i < rs->len[r]; i++) rs->pos2rid[rs->off[r] + i] = (int)r;
  // 10-mer index
  const int K = 10, NK = 1 << 20;
  rs->kstart.assign(NK + 1, 0);
  int n2 = 2 * L;
  auto key_at = [&](int q) {
    uint32_t k = 0;
    for (int j = 0; j < K; j++) k = (k << 2) | rs->T[q + j];
    return k;
  };
  for (int q = 0; q + K <= n2; q++) rs->kstart[key_at(q) + 1]++;
  for (int i = 0; i < NK; i++) rs->kstart[i + 1] += rs->kstart[i];
  rs->kpos.resize(std::max(0, n2 - K + 1));
  std::vector<int> fill(rs->kstart.begin(), rs->kstart.end() - 1);
  for (int q = 0; q + K <= n2; q++) rs->kpos[fill[key_at(q)]++] = q;
  return rs;
}

// bns_intv2rid: record id of a seed's reference span in the 2-strand text, or -1 when it bridges two
// records or the forward/reverse boundary (last bullets).
inline int span_rid(const RefSet& rs, int rb, int re, int* strand) {
  const int L = rs.L;
  if (rb < L && re > L) return -1;
  int b = rb < L ? rb : 2 * L - 1 - rb;
  int e = (re - 1) < L ? (re - 1) : 2 * L - 1 - (re - 1);
  int r1 = rs.pos2rid[b], r2 = rs.pos2rid[e];
  *strand = rb >= L;
  return r1 == r2 ? r1 : -1;
}

// ----------------------------------------------------------------------------------------------
// seeding: which (record, strand) pairs does bwa-mem extend?                 SURVEY A3
// ----------------------------------------------------------------------------------------------
struct Run { int qs, qe, ts; };   // maximal exact-match run: read[qs,qe) == T[ts, ts+qe-qs)

struct SeedStats { long runs = 0, smems = 0, reseeds = 0, tiles = 0, over_max_occ = 0; };

void find_runs(const RefSet& rs, const uint8_t* q, int n, std::vector<Run>& runs) {
  runs.clear();
  const int K = 10, n2 = 2 * rs.L;
  if (n < K) return;
  uint32_t key = 0;
  int valid = 0;
  for (int i = 0; i < n; i++) {
    if (q[i] > 3) { valid = 0; key = 0; continue; }
    key = ((key << 2) | q[i]) & 0xFFFFF;
    if (++valid < K) continue;
    int s = i - K + 1;
    for (int h = rs.kstart[key]; h < rs.kstart[key + 1]; h++) {
      int p = rs.kpos[h];
      if (s > 0 && p > 0 && q[s - 1] < 4 && q[s - 1] == rs.T[p - 1]) continue;   // not left-maximal
      int e = K;
      while (s + e < n && p + e < n2 && q[s + e] < 4 && q[s + e] == rs.T[p + e]) e++;
      runs.push_back({s, s + e, p});
    }
  }
}

// number of occurrences in T of read[s,e) (e-s >= 10): every occurrence lies inside exactly one run.
inline int occ_of(const std::vector<Run>& runs, int s, int e) {
  int c = 0;
  for (const Run& r : runs) c += (r.qs <= s && r.qe >= e);
  return c;
}

// Marks cand[rid*2+strand] for every surviving seed.  Seeds = occurrences of
//   pass 1: SMEMs (len >= 10),
//   pass 2: re-seeding inside SMEMs with len >= 15 and <= 10 occurrences (bwt_smem1 at the middle with
//           min_intv = occ+1),
//   pass 3: the LAST-like tiling (bwt_seed_strategy1, min_len 10 => 11-mers, max_intv 20).
void seed_candidates(const RefSet& rs, const uint8_t* q, int n, std::vector<uint8_t>& cand,
                     std::vector<Run>& runs, SeedStats* st) {
  std::fill(cand.begin(), cand.end(), 0);
  find_runs(rs, q, n, runs);
  if (runs.empty()) return;
  if (st) st->runs += (long)runs.size();
  auto mark = [&](int rb, int len) {
    int strand, rid = span_rid(rs, rb, rb + len, &strand);
    if (rid >= 0) cand[rid * 2 + strand] = 1;
  };
  // ---- pass 1: SMEM = query interval of a run not strictly contained in another run's interval
  std::vector<std::pair<int, int>> smem;   // distinct intervals
  for (const Run& r : runs) {
    bool contained = false;
    for (const Run& o : runs)
      if (o.qs <= r.qs && o.qe >= r.qe && (o.qs < r.qs || o.qe > r.qe)) { contained = true; break; }
    if (contained) continue;
    mark(r.ts, r.qe - r.qs);
    std::pair<int, int> iv(r.qs, r.qe);
    if (std::find(smem.begin(), smem.end(), iv) == smem.end()) smem.push_back(iv);
  }
  if (st) st->smems += (long)smem.size();
  // ---- pass 2: re-seeding
  for (auto& iv : smem) {
    int a = iv.first, b = iv.second;
    int occ = 0;
    for (const Run& r : runs) occ += (r.qs == a && r.qe == b);
    if (st && occ > 500) st->over_max_occ++;
    if (b - a < 15 || occ > 10) continue;
    int mid = (a + b) >> 1, m = occ + 1;
    // runs covering mid; maximal [s,e) containing mid with >= m covering runs:
    std::vector<const Run*> cov;
    for (const Run& r : runs) if (r.qs <= mid && r.qe > mid) cov.push_back(&r);
    if ((int)cov.size() < m) continue;
    std::vector<int> starts;
    for (auto* r : cov) starts.push_back(r->qs);
    std::sort(starts.begin(), starts.end());
    starts.erase(std::unique(starts.begin(), starts.end()), starts.end());
    int prev_e = -1;
    for (int s : starts) {          // ascending s; e*(s) = m-th largest end among runs with qs <= s
      std::vector<int> ends;
      for (auto* r : cov) if (r->qs <= s) ends.push_back(r->qe);
      if ((int)ends.size() < m) continue;
      std::sort(ends.begin(), ends.end(), std::greater<int>());
      int e = ends[m - 1];
      if (e <= prev_e) continue;    // contained in the interval found for a smaller s
      prev_e = e;
      if (e - s < 10) continue;
      if (st) st->reseeds++;
      for (auto* r : cov)
        if (r->qs <= s && r->qe >= e) mark(r->ts + (s - r->qs), e - s);
    }
  }
  // ---- pass 3: tiles
  int x = 0;
  while (x < n) {
    if (q[x] > 3) { x++; continue; }
    int i = x + 1;
    bool emitted = false;
    for (; i < n; i++) {
      if (q[i] > 3) break;                       // return i+1 without a seed
      if (i - x >= 10) {
        int o = occ_of(runs, x, i + 1);
        if (o < 20) {
          if (o > 0) {
            if (st) st->tiles++;
            for (const Run& r : runs)
              if (r.qs <= x && r.qe >= i + 1) mark(r.ts + (x - r.qs), i + 1 - x);
          }
          emitted = true;
          break;
        }
      }
    }
    if (i >= n && !emitted) break;               // ran off the read end: return len
    x = i + 1;
  }
}

// ----------------------------------------------------------------------------------------------
// alignment: local affine SW of the strand-normalised read against ONE record     SURVEY A4, A6
//   match +1, mismatch -2, gap of length k = -(6+k), anything vs N = -1.
//   End cell = first maximum in row-major order (read row outer, record column inner).
// ----------------------------------------------------------------------------------------------
constexpr int GAP_O = 7;   // cost of the first gap base (6 open + 1 extend)
constexpr int GAP_E = 1;

inline int sub_score(uint8_t a, uint8_t b) { return (a > 3 || b > 3) ? -1 : (a == b ? 1 : -2); }

struct SwEnd { int score, qi, tj; };   // qi/tj = 0-based cell of the maximum (inclusive)

SwEnd sw_score(const uint8_t* q, int m, const uint8_t* t, int n) {
  std::vector<int> H(n + 1, 0), E(n + 1, 0);
  SwEnd best{0, -1, -1};
  for (int i = 0; i < m; i++) {
    int hdiag = 0, f = 0, hleft = 0;
    for (int j = 0; j < n; j++) {
      int e = std::max(H[j + 1] - GAP_O, E[j + 1] - GAP_E);   // vertical: consumes read base (I)
      f = std::max(hleft - GAP_O, f - GAP_E);                 // horizontal: consumes record base (D)
      int h = std::max(std::max(0, hdiag + sub_score(q[i], t[j])), std::max(e, f));
      hdiag = H[j + 1];
      H[j + 1] = h; E[j + 1] = e; hleft = h;
      if (h > best.score) best = {h, i, j};
    }
  }
  return best;
}

struct AlnFacts {
  int qb, qe, rb, re;      // half-open, SAM orientation, record-local reference coordinates
  int first_m;             // length of the first M block
  int nm, mi;              // NM (mismatches + gap bases), sum of M and I lengths
  int n_gap_ops;
};

// Traceback from the end cell.  Per-cell preference M, then D (record base skipped), then I — the
// priority of ksw_global2 that bwa uses to lay out its CIGAR; a gap prefers "open" over "extend" on a
// tie; stops at the first H == 0 (shortest alignment).
AlnFacts sw_traceback(const uint8_t* q, int m, const uint8_t* t, int n, const SwEnd& end) {
  const int W = n + 1;
  std::vector<int> H((size_t)(m + 1) * W, 0), E((size_t)(m + 1) * W, 0), F((size_t)(m + 1) * W, 0);
  const int NEG = -1000000;
  for (int j = 0; j <= n; j++) E[j] = F[j] = NEG;
  for (int i = 1; i <= end.qi + 1; i++) {
    E[(size_t)i * W] = F[(size_t)i * W] = NEG;
    for (int j = 1; j <= n; j++) {
      size_t c = (size_t)i * W + j;
      int e = std::max(H[c - W] - GAP_O, E[c - W] - GAP_E);
      int f = std::max(H[c - 1] - GAP_O, F[c - 1] - GAP_E);
      int h = std::max(std::max(0, H[c - W - 1] + sub_score(q[i - 1], t[j - 1])), std::max(e, f));
      H[c] = h; E[c] = e; F[c] = f;
    }
  }
  AlnFacts a{};
  a.qe = end.qi + 1; a.re = end.tj + 1;
  int i = end.qi + 1, j = end.tj + 1, state = 0;   // 0=H, 1=D(F), 2=I(E)
  std::vector<char> ops;
  while (i > 0 && j > 0) {
    size_t c = (size_t)i * W + j;
    if (state == 0) {
      if (H[c] == 0) break;
      if (H[c] == H[c - W - 1] + sub_score(q[i - 1], t[j - 1])) {
        ops.push_back(q[i - 1] < 4 && q[i - 1] == t[j - 1] ? '=' : 'X'); i--; j--;
      } else if (H[c] == F[c]) state = 1;
      else state = 2;
    } else if (state == 1) {
      ops.push_back('D');
      if (F[c] == H[c - 1] - GAP_O) state = 0;
      j--;
    } else {
      ops.push_back('I');
      if (E[c] == H[c - W] - GAP_O) state = 0;
      i--;
    }
  }
  a.qb = i; a.rb = j;
  std::reverse(ops.begin(), ops.end());
  size_t k = 0;
  while (k < ops.size() && (ops[k] == '=' || ops[k] == 'X')) k++;
  a.first_m = (int)k;
  char prev = 0;
  for (char o : ops) {
    if (o == 'X' || o == 'D' || o == 'I') a.nm++;
    if (o != 'D') a.mi++;
    if ((o == 'D' || o == 'I') && o != prev) a.n_gap_ops++;
    prev = o;
  }
  return a;
}

// One SAM-visible alignment record (the winner of one read against one reference set).
struct AlnRec {
  int32_t ordinal = -1;
  int32_t rid = -1;
  int32_t strand = 0;
  int32_t mapped = 0;
  int32_t AS = 0;
  int32_t qb = 0;        // clip5 in SAM orientation; start = qb+1 (GetStartCigar, Jtool.jl:96)
  int32_t m_end = 0;     // term = end of first M block, 1-based (GetEndCigar, Jtool.jl:108)
  int32_t rb = 0;        // POS-1, record-local
  int32_t qe = 0, re = 0;
  int32_t nm = 0, mi = 0;
  int32_t ncand = 0;     // candidates scored
  int32_t ntied = 0;     // candidates sharing the top score
  int64_t cells = 0;     // sum len(read)*len(record) over scored candidates
};

struct Read {
  std::string name;      // trimmed QNAME
  std::string seq, qual; // as in the file
};

void revcomp_codes(const std::vector<uint8_t>& in, std::vector<uint8_t>& out) {
  size_t n = in.size();
  out.resize(n);
  for (size_t i = 0; i < n; i++) out[n - 1 - i] = in[i] > 3 ? 4 : (uint8_t)(3 - in[i]);
}

AlnRec align_read(const RefSet& rs, const std::string& seq, int64_t ordinal, int min_score,
                  std::vector<uint8_t>& cand, std::vector<Run>& runs, SeedStats* st,
                  std::vector<int32_t>* cand_dump /* (rid,strand,score) triples, optional */) {
  AlnRec rec;
  rec.ordinal = (int32_t)ordinal;
  int n = (int)seq.size();
  std::vector<uint8_t> q(n), qrc;
  for (int i = 0; i < n; i++) q[i] = nt_code(seq[i]);
  seed_candidates(rs, q.data(), n, cand, runs, st);
  revcomp_codes(q, qrc);
  struct C { int rid, strand, score, qi, tj; long tpos; };
  std::vector<C> cs;
  int nrec = (int)rs.names.size();
  for (int r = 0; r < nrec; r++)
    for (int s = 0; s < 2; s++) {
      if (!cand[r * 2 + s]) continue;
      const uint8_t* qq = s ? qrc.data() : q.data();
      SwEnd e = sw_score(qq, n, &rs.T[rs.off[r]], rs.len[r]);
      // position of the record on the 2-strand text: forward records ascending, reverse descending
      long tpos = s ? (2L * rs.L - (rs.off[r] + rs.len[r])) : rs.off[r];
      cs.push_back({r, s, e.score, e.qi, e.tj, tpos});
      rec.cells += (int64_t)n * rs.len[r];
      if (cand_dump) { cand_dump->push_back(r); cand_dump->push_back(s); cand_dump->push_back(e.score); }
    }
  rec.ncand = (int)cs.size();
  if (cs.empty()) return rec;
  int mx = 0;
  for (auto& c : cs) mx = std::max(mx, c.score);
  if (mx < min_score) { rec.AS = mx; return rec; }
  std::vector<C> tied;
  for (auto& c : cs) if (c.score == mx) tied.push_back(c);
  std::sort(tied.begin(), tied.end(), [](const C& a, const C& b) { return a.tpos < b.tpos; });
  size_t win = 0;
  uint64_t best_h = ~0ULL;
  for (size_t r = 0; r < tied.size(); r++) {
    uint64_t h = hash_64((uint64_t)ordinal + r);
    if (h < best_h) { best_h = h; win = r; }
  }
  const C& w = tied[win];
  const uint8_t* qq = w.strand ? qrc.data() : q.data();
  AlnFacts a = sw_traceback(qq, n, &rs.T[rs.off[w.rid]], rs.len[w.rid], SwEnd{w.score, w.qi, w.tj});
  rec.mapped = 1; rec.rid = w.rid; rec.strand = w.strand; rec.AS = mx; rec.ntied = (int)tied.size();
  rec.qb = a.qb; rec.m_end = a.qb + a.first_m; rec.rb = a.rb; rec.qe = a.qe; rec.re = a.re;
  rec.nm = a.nm; rec.mi = a.mi;
  return rec;
}

// ----------------------------------------------------------------------------------------------
// FASTQ input (bwa: QNAME = header up to whitespace, trailing /<digit> trimmed)        SURVEY A1
// ----------------------------------------------------------------------------------------------
bool load_fastq(const std::string& path, std::vector<Read>& reads, int64_t max_reads) {
  std::string txt;
  if (!read_file(path, txt)) return false;
  size_t p = 0;
  auto next_line = [&](std::string& out) {
    if (p >= txt.size()) return false;
    size_t e = txt.find('\n', p);
    if (e == std::string::npos) e = txt.size();
    out.assign(txt, p, e - p);
    if (!out.empty() && out.back() == '\r') out.pop_back();
    p = e + 1;
    return true;
  };
  std::string h, s, plus, ql;
  while (next_line(h)) {
    if (h.empty()) continue;
    if (h[0] != '@' && h[0] != '>') return false;
    Read r;
    size_t k = 1;
    while (k < h.size() && !isspace((unsigned char)h[k])) k++;
    r.name = h.substr(1, k - 1);
    size_t l = r.name.size();
    if (l > 2 && r.name[l - 2] == '/' && isdigit((unsigned char)r.name[l - 1])) r.name.resize(l - 2);
    if (!next_line(s)) return false;
    r.seq = s;
    if (h[0] == '@') {
      if (!next_line(plus) || !next_line(ql)) return false;
      r.qual = ql;
    } else {
      r.qual = std::string(s.size(), 'I');
    }
    reads.push_back(std::move(r));
    if (max_reads > 0 && (int64_t)reads.size() >= max_reads) break;
  }
  return true;
}

std::string sam_seq(const std::string& seq, int strand) {
  std::string o(seq.size(), 'N');
  size_t n = seq.size();
  if (!strand) { for (size_t i = 0; i < n; i++) o[i] = NT[nt_code(seq[i])]; }
  else { for (size_t i = 0; i < n; i++) { uint8_t c = nt_code(seq[i]); o[n - 1 - i] = c > 3 ? 'N' : NT[3 - c]; } }
  return o;
}
std::string sam_qual(const std::string& q, int strand) {
  if (!strand) return q;
  return std::string(q.rbegin(), q.rend());
}

// ----------------------------------------------------------------------------------------------
// motif "regex" compiler: every refg motif (reference.jl:13-116) is a union of fixed-length residue
// patterns; expand to alternatives of per-position 32-bit residue masks.            SURVEY A9
// ----------------------------------------------------------------------------------------------
inline int aa_bit(char c) { return (c >= 'A' && c <= 'Z') ? (c - 'A') : (c == '*' ? 26 : -1); }

struct Motif {
  int len = 0;
  std::vector<std::vector<uint32_t>> alts;
  bool match_at(const char* s, int n, int p) const {
    if (len == 0 || p + len > n) return false;
    for (auto& a : alts) {
      bool ok = true;
      for (int k = 0; k < len; k++) {
        int b = aa_bit(s[p + k]);
        if (b < 0 || !((a[k] >> b) & 1)) { ok = false; break; }
      }
      if (ok) return true;
    }
    return false;
  }
  // PCRE eachmatch: leftmost, non-overlapping; returns 1-based offsets
  void eachmatch(const char* s, int n, std::vector<int>& out) const {
    out.clear();
    int p = 0;
    while (p + len <= n && len > 0) {
      if (match_at(s, n, p)) { out.push_back(p + 1); p += len; } else p++;
    }
  }
};

using AltList = std::vector<std::vector<uint32_t>>;
struct RegexParser {
  const std::string& s; size_t p = 0; bool ok = true;
  explicit RegexParser(const std::string& str) : s(str) {}
  static AltList concat(const AltList& a, const AltList& b) {
    AltList r;
    for (auto& x : a) for (auto& y : b) { auto z = x; z.insert(z.end(), y.begin(), y.end()); r.push_back(z); }
    return r;
  }
  AltList parse_alt() {                 // alt := seq ('|' seq)*
    AltList r = parse_seq();
    while (p < s.size() && s[p] == '|') { p++; AltList t = parse_seq(); r.insert(r.end(), t.begin(), t.end()); }
    return r;
  }
  AltList parse_seq() {                 // seq := atom*
    AltList r{{}};
    while (p < s.size() && s[p] != '|' && s[p] != ')') {
      AltList a = parse_atom();
      // quantifier {n}
      if (p < s.size() && s[p] == '{') {
        size_t e = s.find('}', p);
        int n = atoi(s.substr(p + 1, e - p - 1).c_str());
        p = e + 1;
        AltList rep{{}};
        for (int i = 0; i < n; i++) rep = concat(rep, a);
        a = rep;
      }
      r = concat(r, a);
    }
    return r;
  }
  AltList parse_atom() {
    char c = s[p];
    if (c == '(') { p++; AltList r = parse_alt(); if (p < s.size() && s[p] == ')') p++; else ok = false; return r; }
    if (c == '[') {
      p++;
      uint32_t m = 0;
      while (p < s.size() && s[p] != ']') {
        if (p + 2 < s.size() && s[p + 1] == '-' && s[p + 2] != ']') {
          for (char x = s[p]; x <= s[p + 2]; x++) { int b = aa_bit(x); if (b >= 0) m |= 1u << b; }
          p += 3;
        } else { int b = aa_bit(s[p]); if (b >= 0) m |= 1u << b; p++; }
      }
      p++;
      return AltList{{m}};
    }
    p++;
    int b = aa_bit(c);
    return AltList{{b >= 0 ? (1u << b) : 0u}};   // characters outside the residue alphabet never match
  }
};

Motif compile_motif(const std::string& re) {
  Motif m;
  RegexParser ps(re);
  AltList a = ps.parse_alt();
  if (a.empty()) return m;
  m.len = (int)a[0].size();
  for (auto& x : a) if ((int)x.size() != m.len) { fprintf(stderr, "[oracle] motif '%s' is not fixed-length\n", re.c_str()); m.len = 0; m.alts.clear(); return m; }
  m.alts = a;
  return m;
}

const char* CODON =  // index = b1*16+b2*4+b3 with A0 C1 G2 T3 (standard code; Jtool.jl:167-232 lists the same table)
    "KNKNTTTTRSRSIIMIQHQHPPPPRRRRLLLLEDEDAAAAGGGGVVVV*Y*YSSSS*CWCLFLF";

inline void translate(const char* nt, int n, int shift1, std::string& aa) {   // shift1 in 1..3
  aa.clear();
  for (int p = shift1 - 1; p + 3 <= n; p += 3) {
    uint8_t a = nt_code(nt[p]), b = nt_code(nt[p + 1]), c = nt_code(nt[p + 2]);
    aa.push_back((a | b | c) > 3 ? 'X' : CODON[a * 16 + b * 4 + c]);
  }
}

struct Config {
  Motif cmotif, fmotif, innerC, innerF;       // refg entries
  Motif hardC, hardF;                         // Jtool.jl:288,294 (single-read finder ignores refg inner motifs)
  int coffset = 0, foffset = 0;
  int min_score = 12;                         // --bowsc
  int kmer = 10000;                           // 10000 = adaptive (catt.jl:156)
  int nthreads = 1;                           // JULIA_NUM_THREADS (record order, catt.jl:228)
  int allow_v = 3, allow_j = 9;               // real_score allowances (catt.jl:100,121)
};

// Extract_Motif2 (Jtool.jl:234-316).  use_refg_inner selects the 7-arg (array) version.
int extract_motif2(const std::string& aa, const Config& cf, bool use_refg_inner, int* r_lo, int* r_hi) {
  int res = 0;
  std::vector<int> Cx, Fx;
  const int n = (int)aa.size();
  cf.cmotif.eachmatch(aa.data(), n, Cx); for (int& x : Cx) x += cf.coffset;
  cf.fmotif.eachmatch(aa.data(), n, Fx); for (int& x : Fx) x += cf.foffset;
  if (Cx.empty() != Fx.empty()) {
    if (Cx.empty()) { (use_refg_inner ? cf.innerC : cf.hardC).eachmatch(aa.data(), n, Cx); }
    if (Fx.empty()) { (use_refg_inner ? cf.innerF : cf.hardF).eachmatch(aa.data(), n, Fx); for (int& x : Fx) x += 2; }
    res = 2;
  }
  std::reverse(Cx.begin(), Cx.end());
  if (!Cx.empty() && !Fx.empty()) {
    for (size_t idx = 0; idx < Cx.size(); idx++) {
      int xc = Cx[idx];
      for (int xf : Fx) {
        int d = xf - xc;
        if (d < 6 || d > 34) continue;
        if (idx + 1 < Cx.size()) { int d2 = xf - Cx[idx + 1]; if (d2 >= 7 && d2 <= 35) continue; }
        bool stop = false;
        for (int k = xc; k <= xf && !stop; k++) stop = (k >= 1 && k <= n && aa[k - 1] == '*');
        if (stop) continue;
        res = 3; *r_lo = (xc - 1) * 3 + 1; *r_hi = xf * 3;
        break;
      }
    }
  }
  return res;
}

// Myread (Jtool.jl:12-19)
struct Myread {
  std::string seq, qual, vs = "None", js = "None", cdr3 = "None";
  bool able = false;
  int64_t ordinal = -1;    // provenance (not in the reference struct)
  int flags = 0;           // 1 = came from V list part read, 2 = J part read, 4 = both-mapped; 8 = qual slice was out of range
};

// finder! (Jtool.jl:355-389)
void finder(Myread& rd, const Config& cf, bool array_version) {
  const int lgt = (int)rd.seq.size();
  std::string aa;
  for (int shift = 1; shift <= 3; shift++) {
    translate(rd.seq.data(), lgt, shift, aa);
    int lo = 1, hi = 2;
    int code = extract_motif2(aa, cf, array_version, &lo, &hi);
    rd.able = rd.able || code > 0;
    if (code == 3) {
      lo += shift - 1; hi += shift - 1;
      rd.cdr3 = rd.seq.substr(lo - 1, hi - lo + 1);
      // rd.qual[lo:hi]; J-side reads carry a qual string shorter than seq (catt.jl:125-126): the reference
      // would throw a BoundsError if the range ran past it.  Pad with 'G' and flag instead.
      std::string qn;
      for (int k = lo; k <= hi; k++) {
        if (k - 1 < (int)rd.qual.size()) qn.push_back(rd.qual[k - 1]); else { qn.push_back('G'); rd.flags |= 8; }
      }
      rd.qual = qn;
      break;
    }
  }
}

// real_score (catt.jl:69-84) on a record; seq = SAM-oriented read, ref = FASTA bases of the record
bool real_score(const AlnRec& a, const std::string& seq, const uint8_t* ref, int reflen, int allowance) {
  int start = a.qb + 1, term = a.m_end;
  int lgt = term - start;
  int r_pos = a.rb + 1;
  int rdlen = (int)seq.size();
  int left_pad = std::min(start, r_pos) - 1;
  int right_pad = std::max(0, std::min(rdlen - term, reflen - 9 - (r_pos + lgt)));
  int n = lgt + 1 + left_pad + right_pad;
  int upper = (left_pad + right_pad) / 3 + allowance;
  int cnt = 0;
  for (int idx = 0; idx < n; idx++) {
    uint8_t c1 = nt_code(seq[start - left_pad - 1 + idx]);
    uint8_t c2 = ref[r_pos - left_pad - 1 + idx];
    if (c1 != c2 && ++cnt > upper) return false;
  }
  return true;
}

inline bool has_n(const std::string& s) { for (char c : s) if (nt_code(c) > 3) return true; return false; }

struct RunInfo {
  int kmer = 0;
  double the_err = 0.0;
  double avg_len = 0.0;
  long n_reads = 0, v_hits = 0, j_hits = 0, v_gapped = 0, j_gapped = 0, v_tied = 0, j_tied = 0;
  long v_final = 0, j_final = 0, share = 0, pair_seq_mismatch = 0, full_pushed = 0;
  long vpart = 0, jpart = 0, direct = 0, left = 0, rescued = 0;
  long cand_v = 0, cand_j = 0;
  int64_t cells_v = 0, cells_j = 0;
  long qual_oob = 0;
  double t_align = 0, t_extract = 0, t_extend = 0;
};

struct RunResult {
  std::vector<Myread> Vpart, Jpart;
  std::vector<AlnRec> vrec, jrec;     // per read, input order
  RunInfo info;
  std::string dump;
};

// k-mer helpers (2-bit, k <= 31)
inline bool kmer_at(const std::string& s, int pos0, int k, uint64_t* out) {
  uint64_t v = 0;
  for (int i = 0; i < k; i++) { uint8_t c = nt_code(s[pos0 + i]); if (c > 3) return false; v = (v << 2) | c; }
  *out = v; return true;
}
inline bool ratio_ok(long x, long y) { return x > y ? (x <= 10 * y) : (10 * x >= y); }   // catt.jl:341

using Pool = std::unordered_map<uint64_t, int>;

inline int pool_get(const Pool& p, uint64_t key) { auto it = p.find(key); return it == p.end() ? 0 : it->second; }

// extend_right! (catt.jl:350-382); neighbours = bw_neighbors(seg) = c + seg[1:k-1] for c in ACGT (sic)
void extend_right(Myread& rd, const Pool& pool, int k) {
  int len = (int)rd.seq.size();
  uint64_t seg;
  if (len < k || !kmer_at(rd.seq, len - k, k, &seg)) return;
  long cur = pool_get(pool, seg);
  std::string add;
  int idx = 1; bool flag = true;
  while (idx < 151 - len && flag) {
    uint64_t cand[4]; long val[4]; int ord[4] = {0, 1, 2, 3};
    for (int c = 0; c < 4; c++) { cand[c] = ((uint64_t)c << (2 * (k - 1))) | (seg >> 2); val[c] = pool_get(pool, cand[c]); }
    std::stable_sort(ord, ord + 4, [&](int a, int b) { return val[a] > val[b]; });
    flag = false;
    for (int t = 0; t < 4; t++) {
      int c = ord[t];
      if (val[c] == 0) continue;
      if (ratio_ok(val[c], cur) && cand[c] != seg) { seg = cand[c]; cur = val[c]; add.push_back(NT[c]); idx++; flag = true; break; }
    }
  }
  rd.seq += add;
  rd.qual += std::string(add.size(), 'G');
}

// extend_left! (catt.jl:391-424); neighbours = fw_neighbors(seg) = seg[2:k] + c (sic)
void extend_left(Myread& rd, const Pool& pool, int k) {
  int len = (int)rd.seq.size();
  uint64_t seg;
  if (len < k || !kmer_at(rd.seq, 0, k, &seg)) return;
  const uint64_t mask = (k == 32) ? ~0ULL : ((1ULL << (2 * k)) - 1);
  long cur = pool_get(pool, seg);
  std::string add;
  int idx = 1; bool flag = true;
  while (idx < 151 - len && flag) {
    uint64_t cand[4]; long val[4]; int ord[4] = {0, 1, 2, 3};
    for (int c = 0; c < 4; c++) { cand[c] = ((seg << 2) | (uint64_t)c) & mask; val[c] = pool_get(pool, cand[c]); }
    std::stable_sort(ord, ord + 4, [&](int a, int b) { return val[a] > val[b]; });
    flag = false;
    for (int t = 0; t < 4; t++) {
      int c = ord[t];
      if (val[c] == 0) continue;
      if (ratio_ok(val[c], cur) && cand[c] != seg) { seg = cand[c]; cur = val[c]; add.push_back(NT[c]); idx++; flag = true; break; }
    }
  }
  std::string r(add.rbegin(), add.rend());
  rd.seq = r + rd.seq;
  rd.qual = std::string(add.size(), 'G') + rd.qual;
}

double now_s() {
#ifdef _OPENMP
  return omp_get_wtime();
#else
  return (double)clock() / CLOCKS_PER_SEC;
#endif
}

// final SAM list of one reference set: AS-sorted, reversed, first record per QNAME   SURVEY A7
std::vector<int> final_sam_order(const std::vector<AlnRec>& recs, const std::vector<Read>& reads, size_t lo, size_t hi) {
  std::vector<int> idx;
  for (size_t i = lo; i < hi; i++) if (recs[i].mapped) idx.push_back((int)i);
  std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) {
    const AlnRec &x = recs[a], &y = recs[b];
    if (x.AS != y.AS) return x.AS < y.AS;
    if (x.rid != y.rid) return x.rid < y.rid;
    if (x.rb != y.rb) return x.rb < y.rb;
    return x.strand < y.strand;
  });
  std::reverse(idx.begin(), idx.end());
  std::unordered_set<std::string> seen;
  std::vector<int> out;
  for (int i : idx) if (seen.insert(reads[i].name).second) out.push_back(i);
  return out;
}

// Julia reads chunk files p = NR % T for p = 0..T-1 (catt.jl:228,256-276)
std::vector<int> chunk_order(const std::vector<int>& fin, int T) {
  if (T <= 1) return fin;
  std::vector<int> out;
  for (int p = 0; p < T; p++)
    for (size_t nr = 1; nr <= fin.size(); nr++)
      if ((int)(nr % T) == p) out.push_back(fin[nr - 1]);
  return out;
}

// `file_start` (nullptr = one file) lists the first read of every input file plus the total: the reference runs
// map2align and the read_alignrs loop once per file of a paired-end sample (catt.jl:134-151, 210-326: bwa numbers the
// reads of each file from 0, the awk dedup / share_name / both-mapped pairing are per file, args["kmer"] is fixed by the
// first file whose mean length reaches 50, the_ERR is overwritten per file) and then catt() once over the joined lists.
RunResult* run_pipeline(const RefSet& V, const RefSet& J, const Config& cf0, const std::vector<Read>& reads,
                        int omp_threads, int want_dump, const std::vector<int64_t>* file_start = nullptr) {
  Config cf = cf0;
  RunResult* R = new RunResult();
  RunInfo& info = R->info;
  const int64_t n = (int64_t)reads.size();
  std::vector<int64_t> fs = file_start ? *file_start : std::vector<int64_t>{0, n};
  info.n_reads = n;
  R->vrec.resize(n); R->jrec.resize(n);
  double t0 = now_s();
  // ---------------- stage 1+2: seeding + SW + selection (map2align, catt.jl:37-56)
#ifdef _OPENMP
  if (omp_threads > 0) omp_set_num_threads(omp_threads);
#endif
  std::vector<int64_t> file_base((size_t)n, 0);
  for (size_t f = 0; f + 1 < fs.size(); f++) for (int64_t i = fs[f]; i < fs[f + 1]; i++) file_base[i] = fs[f];
  #pragma omp parallel
  {
    std::vector<uint8_t> candV(V.names.size() * 2), candJ(J.names.size() * 2);
    std::vector<Run> runs;
    #pragma omp for schedule(dynamic, 64)
    for (int64_t i = 0; i < n; i++) {
      R->vrec[i] = align_read(V, reads[i].seq, i - file_base[i], cf.min_score, candV, runs, nullptr, nullptr);
      R->jrec[i] = align_read(J, reads[i].seq, i - file_base[i], cf.min_score, candJ, runs, nullptr, nullptr);
    }
  }
  for (int64_t i = 0; i < n; i++) {
    const AlnRec &a = R->vrec[i], &b = R->jrec[i];
    info.cand_v += a.ncand; info.cand_j += b.ncand; info.cells_v += a.cells; info.cells_j += b.cells;
    if (a.mapped) { info.v_hits++; info.v_gapped += (a.m_end != a.qe); info.v_tied += a.ntied > 1; }
    if (b.mapped) { info.j_hits++; info.j_gapped += (b.m_end != b.qe); info.j_tied += b.ntied > 1; }
  }
  double t1 = now_s();
  info.t_align = t1 - t0;
  // ---------------- stage 3a: read_alignrs (catt.jl:186-339), once per input file
  for (size_t f = 0; f + 1 < fs.size(); f++) {
  std::vector<int> vfin = final_sam_order(R->vrec, reads, (size_t)fs[f], (size_t)fs[f + 1]);
  std::vector<int> jfin = final_sam_order(R->jrec, reads, (size_t)fs[f], (size_t)fs[f + 1]);
  info.v_final += (long)vfin.size(); info.j_final += (long)jfin.size();
  // get_kmer_fromsam / get_err_fromsam (catt.jl:153-184) on the final V list
  {
    double tot = 0; long nm = 0, mi = 0;
    for (int i : vfin) { tot += (double)reads[i].seq.size(); nm += R->vrec[i].nm; mi += R->vrec[i].mi; }
    double avg = vfin.empty() ? 0.0 : tot / (double)vfin.size();
    char buf[64];
    snprintf(buf, sizeof buf, "%.0f", avg);          // samtools stats "average length" is printed %.0f
    double avg_r = atof(buf);
    info.avg_len = avg_r;
    if (cf.kmer == 10000) {
      if (avg_r >= 50) cf.kmer = 15;
      if (avg_r >= 75) cf.kmer = 19;
      if (avg_r >= 100) cf.kmer = 23;
      if (avg_r >= 150) cf.kmer = 25;
    }
    float er = mi ? (float)nm / (float)mi : 0.0f;      // samtools stats: (float)mismatches / bases mapped (cigar), "%e"
    snprintf(buf, sizeof buf, "%e", (double)er);
    info.the_err = atof(buf);
    if (!(info.the_err == info.the_err)) info.the_err = 0.0;
  }
  info.kmer = cf.kmer;
  const int k = cf.kmer;
  std::unordered_set<std::string> vnames, share;
  for (int i : vfin) vnames.insert(reads[i].name);
  for (int i : jfin) if (vnames.count(reads[i].name)) share.insert(reads[i].name);
  info.share += (long)share.size();

  struct Full { std::string name; int idx; int which; };
  std::vector<Full> full;
  for (int which = 0; which < 2; which++) {
    const RefSet& rs = which ? J : V;
    const std::vector<AlnRec>& recs = which ? R->jrec : R->vrec;
    std::vector<Myread>& col = which ? R->Jpart : R->Vpart;
    for (int i : chunk_order(which ? jfin : vfin, cf.nthreads)) {
      const AlnRec& a = recs[i];
      if (share.count(reads[i].name)) { full.push_back({reads[i].name, i, which}); continue; }
      std::string tseq = sam_seq(reads[i].seq, a.strand), tq = sam_qual(reads[i].qual, a.strand);
      const int rdlen = (int)tseq.size(), start = a.qb + 1, r_pos = a.rb + 1, r_lgt = rs.len[a.rid];
      const uint8_t* ref = &rs.fa[rs.off[a.rid]];
      Myread rd;
      rd.ordinal = i;
      if (!which) {                                                   // assignV (catt.jl:86-106)
        if ((r_lgt - r_pos) - (rdlen - start) > -10) continue;
        if (!real_score(a, tseq, ref, r_lgt, cf.allow_v)) continue;
        rd.seq = tseq.substr(start - 1); rd.qual = tq.substr(std::min((size_t)(start - 1), tq.size()));
        rd.vs = rs.names[a.rid]; rd.flags = 1;
      } else {                                                        // assignJ (catt.jl:108-132)
        if (start - 1 < 5) continue;
        if (!real_score(a, tseq, ref, r_lgt, cf.allow_j)) continue;
        rd.seq = tseq.substr(0, std::min(rdlen, start + r_lgt));
        rd.qual = tq.substr(0, start - 1) + std::string(r_lgt - r_pos, 'G');
        rd.js = rs.names[a.rid]; rd.flags = 2;
      }
      finder(rd, cf, false);
      if (!rd.able) continue;
      if (!((int)rd.seq.size() > k) || has_n(rd.seq)) continue;       // catt.jl:279
      col.push_back(std::move(rd));
    }
  }
  // both-mapped reads (catt.jl:288-319): stable sort by name keeps (V record, J record) order
  std::stable_sort(full.begin(), full.end(), [](const Full& a, const Full& b) { return a.name < b.name; });
  for (size_t g = 0; g < full.size();) {
    size_t h = g;
    while (h < full.size() && full[h].name == full[g].name) h++;
    if (h - g == 2) {
      const Full &p1 = full[g], &p2 = full[g + 1];
      const AlnRec& a1 = (p1.which ? R->jrec : R->vrec)[p1.idx];
      const AlnRec& a2 = (p2.which ? R->jrec : R->vrec)[p2.idx];
      std::string s1 = sam_seq(reads[p1.idx].seq, a1.strand), s2 = sam_seq(reads[p2.idx].seq, a2.strand);
      if (s1 == s2) {
        int s = a1.qb + 1, t = a2.m_end;
        if (t - s + 1 >= k) {
          std::string fin = s1.substr(s - 1, t - s + 1);
          if (!has_n(fin)) {
            Myread rd;
            rd.ordinal = p1.idx; rd.flags = 4;
            rd.seq = fin; rd.qual = sam_qual(reads[p1.idx].qual, a1.strand).substr(s - 1, t - s + 1);
            rd.vs = (p1.which ? J : V).names[a1.rid]; rd.js = (p2.which ? J : V).names[a2.rid];
            finder(rd, cf, false);
            if (rd.able) { R->Vpart.push_back(std::move(rd)); info.full_pushed++; }
          }
        }
      } else info.pair_seq_mismatch++;
    }
    g = h;
  }
  }  // per input file
  const int k = cf.kmer;
  double t2 = now_s();
  info.t_extract = t2 - t1;
  info.vpart = (long)R->Vpart.size(); info.jpart = (long)R->Jpart.size();
  // ---------------- stage 3b/c: catt() pool + extension (catt.jl:499-577)
  std::vector<Myread*> pV, pJ;
  for (auto& r : R->Vpart) { if (r.cdr3 != "None") info.direct++; else pV.push_back(&r); }
  for (auto& r : R->Jpart) { if (r.cdr3 != "None") info.direct++; else pJ.push_back(&r); }
  info.left = (long)(pV.size() + pJ.size());
  if (k <= 31) {
    Pool pool;
    auto add_frag = [&](const std::string& frag) {            // convert2Dict: res[kmer] = position (catt.jl:426-432)
      for (int p = 0; p + k <= (int)frag.size(); p++) { uint64_t key; if (kmer_at(frag, p, k, &key)) pool[key] = p + 1; }
    };
    for (auto& r : R->Vpart) { int l = (int)r.seq.size(); int f = std::min(l, k + 10); add_frag(r.seq.substr(l - f)); }   // last_serveral
    for (auto& r : R->Jpart) { int l = (int)r.seq.size(); int f = std::min(l, k + 10); add_frag(r.seq.substr(0, f)); }    // first_serveral
    uint64_t gk = 0; for (int i = 0; i < k; i++) gk = (gk << 2) | 2;
    pool[gk] = 0;                                                 // catt.jl:545-548
    for (Myread* r : pV) { extend_right(*r, pool, k); finder(*r, cf, true); if (r->cdr3 != "None") info.rescued++; }
    for (Myread* r : pJ) { extend_left(*r, pool, k); finder(*r, cf, true); if (r->cdr3 != "None") info.rescued++; }
  }
  for (auto& r : R->Vpart) info.qual_oob += (r.flags & 8) != 0;
  for (auto& r : R->Jpart) info.qual_oob += (r.flags & 8) != 0;
  info.t_extend = now_s() - t2;
  (void)want_dump;
  return R;
}

// selBestMatchD (Jtool.jl:40-52): global affine, match 5, mismatch -3, gap(k) = -(4+k)
int global_affine(const std::string& a, const std::string& b) {
  const int m = (int)a.size(), n = (int)b.size(), NEG = -1000000;
  std::vector<int> H(n + 1), E(n + 1, NEG);
  H[0] = 0;
  for (int j = 1; j <= n; j++) H[j] = -(4 + j);
  for (int i = 1; i <= m; i++) {
    int hdiag = H[0], f = NEG;
    H[0] = -(4 + i);
    for (int j = 1; j <= n; j++) {
      int e = std::max(H[j] - 5, E[j] - 1);
      f = std::max(H[j - 1] - 5, f - 1);
      int s = (toupper(a[i - 1]) == toupper(b[j - 1])) ? 5 : -3;
      int h = std::max(hdiag + s, std::max(e, f));
      hdiag = H[j]; H[j] = h; E[j] = e;
    }
  }
  return H[n];
}

}  // namespace

// ================================================================================================
// C interface (ctypes)
// ================================================================================================
extern "C" {

void* co_refset_load(const char* fasta, int use_pac) { return load_refset(fasta, use_pac); }
void co_refset_free(void* h) { delete (RefSet*)h; }
int co_refset_nrec(void* h) { return (int)((RefSet*)h)->names.size(); }
int co_refset_len(void* h) { return ((RefSet*)h)->L; }
const char* co_refset_name(void* h, int i) { return ((RefSet*)h)->names[i].c_str(); }
int co_refset_reclen(void* h, int i) { return ((RefSet*)h)->len[i]; }
int co_refset_recoff(void* h, int i) { return ((RefSet*)h)->off[i]; }
int co_refset_pac_stats(void* h, int* n_from_pac, int* mismatch) {
  *n_from_pac = ((RefSet*)h)->n_from_pac; *mismatch = ((RefSet*)h)->pac_mismatch; return 0;
}
// alignment bases of the forward text (0..3), L bytes
void co_refset_bases(void* h, uint8_t* out) { RefSet* r = (RefSet*)h; memcpy(out, r->T.data(), r->L); }

struct co_config {
  const char *cmotif, *fmotif, *innerC, *innerF;
  int coffset, foffset, min_score, kmer, nthreads, allow_v, allow_j;
};

static Config make_config(const co_config* c) {
  Config cf;
  cf.cmotif = compile_motif(c->cmotif); cf.fmotif = compile_motif(c->fmotif);
  cf.innerC = compile_motif(c->innerC); cf.innerF = compile_motif(c->innerF);
  cf.hardC = compile_motif("(CAS|CSA|CAW|CAT|CSV|CAI|CAR|CAG|CSG)");
  cf.hardF = compile_motif("(QYF|QFF|AFF|LFF|YTF|QHF|LHF|IYF|LTF)");
  cf.coffset = c->coffset; cf.foffset = c->foffset; cf.min_score = c->min_score; cf.kmer = c->kmer;
  cf.nthreads = c->nthreads < 1 ? 1 : c->nthreads; cf.allow_v = c->allow_v; cf.allow_j = c->allow_j;
  return cf;
}

// Align n reads (NUL-terminated strings) against one reference set.  out: n AlnRec-shaped rows of 16 int32
// [ordinal,rid,strand,mapped,AS,qb,m_end,rb,qe,re,nm,mi,ncand,ntied,cells_lo,cells_hi].
// cand_out (optional): for each read up to max_cand (rid,strand,score) triples, count in ncand.
int co_align(void* refset, int64_t n, const char* const* seqs, int64_t first_ordinal, int min_score,
             int32_t* out, int32_t* cand_out, int max_cand, int omp_threads) {
  const RefSet& rs = *(RefSet*)refset;
#ifdef _OPENMP
  if (omp_threads > 0) omp_set_num_threads(omp_threads);
#endif
  #pragma omp parallel
  {
    std::vector<uint8_t> cand(rs.names.size() * 2);
    std::vector<Run> runs;
    std::vector<int32_t> dump;
    #pragma omp for schedule(dynamic, 64)
    for (int64_t i = 0; i < n; i++) {
      dump.clear();
      AlnRec a = align_read(rs, seqs[i], first_ordinal + i, min_score, cand, runs, nullptr, cand_out ? &dump : nullptr);
      int32_t* o = out + i * 16;
      o[0] = a.ordinal; o[1] = a.rid; o[2] = a.strand; o[3] = a.mapped; o[4] = a.AS; o[5] = a.qb; o[6] = a.m_end; o[7] = a.rb;
      o[8] = a.qe; o[9] = a.re; o[10] = a.nm; o[11] = a.mi; o[12] = a.ncand; o[13] = a.ntied;
      o[14] = (int32_t)(a.cells & 0xffffffff); o[15] = (int32_t)(a.cells >> 32);
      if (cand_out) {
        int32_t* c = cand_out + i * (int64_t)max_cand * 3;
        for (int k = 0; k < max_cand * 3; k++) c[k] = -1;
        for (size_t k = 0; k < dump.size() && (int)k < max_cand * 3; k++) c[k] = dump[k];
      }
    }
  }
  return 0;
}

// candidates only (bitmap of nrec*2 bytes per read) — for seeding parity tests
int co_seed(void* refset, int64_t n, const char* const* seqs, uint8_t* cand_out) {
  const RefSet& rs = *(RefSet*)refset;
  size_t w = rs.names.size() * 2;
  #pragma omp parallel
  {
    std::vector<uint8_t> cand(w);
    std::vector<Run> runs;
    std::vector<uint8_t> q;
    #pragma omp for schedule(dynamic, 64)
    for (int64_t i = 0; i < n; i++) {
      int len = (int)strlen(seqs[i]);
      q.resize(len);
      for (int k = 0; k < len; k++) q[k] = nt_code(seqs[i][k]);
      seed_candidates(rs, q.data(), len, cand, runs, nullptr);
      memcpy(cand_out + i * w, cand.data(), w);
    }
  }
  return 0;
}

// plain SW score of one pair (codes as ASCII), returns score; end cell in *qi,*tj
int co_sw(const char* q, const char* t, int* qi, int* tj) {
  std::vector<uint8_t> a(strlen(q)), b(strlen(t));
  for (size_t i = 0; i < a.size(); i++) a[i] = nt_code(q[i]);
  for (size_t i = 0; i < b.size(); i++) b[i] = nt_code(t[i]);
  SwEnd e = sw_score(a.data(), (int)a.size(), b.data(), (int)b.size());
  if (qi) *qi = e.qi;
  if (tj) *tj = e.tj;
  return e.score;
}

// full traceback facts for one pair: out[0..7] = qb,qe,rb,re,first_m,nm,mi,n_gap_ops; returns score
int co_sw_facts(const char* q, const char* t, int32_t* out) {
  std::vector<uint8_t> a(strlen(q)), b(strlen(t));
  for (size_t i = 0; i < a.size(); i++) a[i] = nt_code(q[i]);
  for (size_t i = 0; i < b.size(); i++) b[i] = nt_code(t[i]);
  SwEnd e = sw_score(a.data(), (int)a.size(), b.data(), (int)b.size());
  if (e.score <= 0) { for (int i = 0; i < 8; i++) out[i] = 0; return 0; }
  AlnFacts f = sw_traceback(a.data(), (int)a.size(), b.data(), (int)b.size(), e);
  out[0] = f.qb; out[1] = f.qe; out[2] = f.rb; out[3] = f.re; out[4] = f.first_m; out[5] = f.nm; out[6] = f.mi; out[7] = f.n_gap_ops;
  return e.score;
}

uint64_t co_hash64(uint64_t k) { return hash_64(k); }

// finder! on one read; returns able | (found<<1); cdr3 range (1-based, inclusive) in lo/hi
int co_finder(const co_config* c, const char* seq, int array_version, int* lo, int* hi) {
  Config cf = make_config(c);
  Myread rd; rd.seq = seq; rd.qual = std::string(rd.seq.size(), 'I');
  finder(rd, cf, array_version != 0);
  *lo = *hi = 0;
  if (rd.cdr3 != "None") { size_t p = rd.seq.find(rd.cdr3); (void)p; }
  // recompute range for reporting
  if (rd.cdr3 != "None") {
    const int lgt = (int)rd.seq.size(); std::string aa;
    for (int shift = 1; shift <= 3; shift++) {
      translate(rd.seq.data(), lgt, shift, aa);
      int l = 1, h = 2;
      if (extract_motif2(aa, cf, array_version != 0, &l, &h) == 3) { *lo = l + shift - 1; *hi = h + shift - 1; break; }
    }
  }
  return (rd.able ? 1 : 0) | (rd.cdr3 != "None" ? 2 : 0);
}

int co_motif_expand(const char* re, int* len) { Motif m = compile_motif(re); *len = m.len; return (int)m.alts.size(); }

int co_global_affine(const char* a, const char* b) { return global_affine(a, b); }

// selBestMatchD: returns index into the D list or -1 ("None")
int co_best_d(const char* cdr3, int nd, const char* const* dseqs) {
  int best = -1, best_score = 0;
  const int ls = (int)strlen(cdr3);
  for (int i = 0; i < nd; i++) {
    int sc = global_affine(cdr3, dseqs[i]);
    if (!(sc + ls - (int)strlen(dseqs[i]) + 4 > 35)) continue;
    if (best < 0 || sc > best_score) { best = i; best_score = sc; }   // stable sort descending: first max wins
  }
  return best;
}

// ---- whole pipeline ------------------------------------------------------------------------------
void* co_run(void* vref, void* jref, const co_config* c, const char* fastq, int64_t max_reads, int omp_threads) {
  std::vector<Read> reads;
  if (!load_fastq(fastq, reads, max_reads)) return nullptr;
  Config cf = make_config(c);
  return run_pipeline(*(RefSet*)vref, *(RefSet*)jref, cf, reads, omp_threads, 0);
}

// same, but reads supplied from memory (names may be NULL => "r<i>")
void* co_run_mem(void* vref, void* jref, const co_config* c, int64_t n, const char* const* names,
                 const char* const* seqs, const char* const* quals, int omp_threads) {
  std::vector<Read> reads((size_t)n);
  for (int64_t i = 0; i < n; i++) {
    Read& r = reads[i];
    r.name = names ? names[i] : ("r" + std::to_string(i));
    size_t l = r.name.size();
    if (l > 2 && r.name[l - 2] == '/' && isdigit((unsigned char)r.name[l - 1])) r.name.resize(l - 2);
    r.seq = seqs[i];
    r.qual = quals ? quals[i] : std::string(r.seq.size(), 'I');
  }
  Config cf = make_config(c);
  return run_pipeline(*(RefSet*)vref, *(RefSet*)jref, cf, reads, omp_threads, 0);
}

// paired-end / multi-file sample from memory: file_start has nfiles+1 entries
void* co_run_mem_files(void* vref, void* jref, const co_config* c, int64_t n, const char* const* names, const char* const* seqs,
                       const char* const* quals, int nfiles, const int64_t* file_start, int omp_threads) {
  std::vector<Read> reads((size_t)n);
  for (int64_t i = 0; i < n; i++) {
    Read& r = reads[i];
    r.name = names[i];
    size_t l = r.name.size();
    if (l > 2 && r.name[l - 2] == '/' && isdigit((unsigned char)r.name[l - 1])) r.name.resize(l - 2);
    r.seq = seqs[i];
    r.qual = quals ? quals[i] : std::string(r.seq.size(), 'I');
  }
  Config cf = make_config(c);
  std::vector<int64_t> fs(file_start, file_start + nfiles + 1);
  return run_pipeline(*(RefSet*)vref, *(RefSet*)jref, cf, reads, omp_threads, 0, &fs);
}

// Synthetic reads of SURVEY §8d config 3 (same algorithm as tests/util.py::synth_read): fixed length, read i drawn from
// splitmix64(seed + i).  seqs/quals: n*length bytes.  Lets the CPU baseline build its own input without the GPU library.
int co_synth_reads(void* vref, void* jref, int64_t n, int64_t first_index, int length, uint64_t seed, double p_err, int with_n,
                   char* seqs, char* quals) {
  const RefSet &V = *(RefSet*)vref, &J = *(RefSet*)jref;
  const int perr = (int)(p_err * 100000);
  #pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < n; r++) {
    uint64_t st = seed + (uint64_t)(first_index + r);
    auto below = [&](uint64_t m) {
      st += 0x9E3779B97F4A7C15ull;
      uint64_t z = st;
      z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
      z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
      return (z ^ (z >> 31)) % m;
    };
    std::string tpl, ins, pre, post;
    const int vi = (int)below(V.names.size()), ji = (int)below(J.names.size());
    const int vlen = V.len[vi] - (int)below(7);
    const int jtrim = (int)below(7);
    const int nins = (int)below(19);
    for (int k = 0; k < nins; k++) ins.push_back(NT[below(4)]);
    for (int k = 0; k < 60; k++) pre.push_back(NT[below(4)]);
    for (int k = 0; k < 120; k++) post.push_back(NT[below(4)]);
    tpl = pre;
    const int v0 = (int)tpl.size();
    for (int k = 0; k < vlen; k++) tpl.push_back(NT[V.T[V.off[vi] + k]]);
    const int v1 = (int)tpl.size();
    tpl += ins;
    const int j0 = (int)tpl.size();
    for (int k = jtrim; k < J.len[ji]; k++) tpl.push_back(NT[J.T[J.off[ji] + k]]);
    const int j1 = (int)tpl.size();
    tpl += post;
    int start = 0;
    for (int t = 0; t < 64; t++) {
      start = (int)below((uint64_t)std::max(1, (int)tpl.size() - length + 1));
      const int en = start + length;
      const int ovv = std::max(0, std::min(en, v1) - std::max(start, v0)), ovj = std::max(0, std::min(en, j1) - std::max(start, j0));
      if (ovv >= 20 || ovj >= 12) break;
    }
    char* s = seqs + r * (int64_t)length;
    char* q = quals + r * (int64_t)length;
    for (int k = 0; k < length; k++) {
      s[k] = start + k < (int)tpl.size() ? tpl[start + k] : 'A';
      q[k] = 'I';
      const int x = (int)below(100000);
      if (x < perr) { s[k] = NT[below(4)]; q[k] = '#'; }
      else if (x < 5000 + perr) q[k] = 'G';
    }
    if (with_n && below(50) == 0) s[below((uint64_t)length)] = 'N';
    if (below(2)) {
      std::string t(s, length);
      for (int k = 0; k < length; k++) { uint8_t c = nt_code(t[length - 1 - k]); s[k] = c > 3 ? 'N' : NT[3 - c]; }
      std::reverse(q, q + length);
    }
  }
  return 0;
}

// whole pipeline on fixed-length reads held in flat buffers (names "s<first_index+i>")
void* co_run_flat(void* vref, void* jref, const co_config* c, int64_t n, int64_t first_index, int length, const char* seqs,
                  const char* quals, int omp_threads) {
  std::vector<Read> reads((size_t)n);
  for (int64_t i = 0; i < n; i++) {
    reads[i].name = "s" + std::to_string(first_index + i);
    reads[i].seq.assign(seqs + i * (int64_t)length, length);
    reads[i].qual.assign(quals + i * (int64_t)length, length);
  }
  Config cf = make_config(c);
  return run_pipeline(*(RefSet*)vref, *(RefSet*)jref, cf, reads, omp_threads, 0);
}

void co_result_free(void* r) { delete (RunResult*)r; }

// info as doubles: see python binding for the field order
int co_result_info(void* r, double* out) {
  const RunInfo& i = ((RunResult*)r)->info;
  double v[] = {(double)i.kmer, i.the_err, i.avg_len, (double)i.n_reads, (double)i.v_hits, (double)i.j_hits, (double)i.v_gapped,
                (double)i.j_gapped, (double)i.v_tied, (double)i.j_tied, (double)i.v_final, (double)i.j_final, (double)i.share,
                (double)i.pair_seq_mismatch, (double)i.full_pushed, (double)i.vpart, (double)i.jpart, (double)i.direct,
                (double)i.left, (double)i.rescued, (double)i.cand_v, (double)i.cand_j, (double)i.cells_v, (double)i.cells_j,
                (double)i.qual_oob, i.t_align, i.t_extract, i.t_extend};
  memcpy(out, v, sizeof v);
  return (int)(sizeof v / sizeof v[0]);
}

int64_t co_result_count(void* r, int which) { RunResult* R = (RunResult*)r; return (int64_t)(which ? R->Jpart : R->Vpart).size(); }

// serialise Vpart (which=0) / Jpart (1) as TSV: seq qual vs js cdr3 able ordinal flags
const char* co_result_dump(void* r, int which) {
  RunResult* R = (RunResult*)r;
  std::string& d = R->dump;
  d.clear();
  for (const Myread& m : (which ? R->Jpart : R->Vpart)) {
    d += m.seq; d += '\t'; d += m.qual; d += '\t'; d += m.vs; d += '\t'; d += m.js; d += '\t'; d += m.cdr3; d += '\t';
    d += m.able ? '1' : '0'; d += '\t'; d += std::to_string(m.ordinal); d += '\t'; d += std::to_string(m.flags); d += '\n';
  }
  return d.c_str();
}

static std::string trim_readno(const char* nm) {          // bwa trim_readno, as in load_fastq
  std::string r(nm);
  const size_t l = r.size();
  if (l > 2 && r[l - 2] == '/' && isdigit((unsigned char)r[l - 1])) r.resize(l - 2);
  return r;
}

// ---- single-stage hooks for the parity tests -------------------------------------------------------------------------------------
// real_score (catt.jl:58-84) alone: raw reads, record id / strand / qb / m_end / rb of a SAM record each; out[i] = 1 when it passes
int co_real_score(void* refset, int64_t n, const char* const* seqs, const int32_t* rid, const int32_t* strand, const int32_t* qb,
                  const int32_t* m_end, const int32_t* rb, int allowance, uint8_t* out) {
  const RefSet& rs = *(RefSet*)refset;
  for (int64_t i = 0; i < n; i++) {
    AlnRec a{};
    a.rid = rid[i]; a.strand = strand[i]; a.qb = qb[i]; a.m_end = m_end[i]; a.rb = rb[i];
    const std::string tseq = sam_seq(seqs[i], a.strand);
    out[i] = real_score(a, tseq, &rs.fa[rs.off[a.rid]], rs.len[a.rid], allowance) ? 1 : 0;
  }
  return 0;
}

// The ordering stage alone (catt.jl:45-48, 221-237, 256-276, 288-297): from per-read records (mapped, AS, rid, rb, strand) and
// QNAMEs to the work items in processing order - [V-only records in chunk order | both-mapped pairs in name order | J-only records
// in chunk order].  items[3 * t] = kind (0 V, 1 J, 2 pair), read index, second read index (the J record's read for a pair).
// counts = {final V records, final J records, V-only items, pairs, J-only items}.  Returns the number of items.
int64_t co_order(int64_t n, const char* const* names, const int32_t* vrec16, const int32_t* jrec16, int nthreads, int32_t* items, int64_t* counts) {
  std::vector<Read> reads((size_t)n);
  std::vector<AlnRec> vr((size_t)n), jr((size_t)n);
  auto fill = [&](const int32_t* src, std::vector<AlnRec>& dst) {
    for (int64_t i = 0; i < n; i++) {
      const int32_t* r = src + i * 16;
      AlnRec a{};
      a.rid = r[1]; a.strand = r[2]; a.mapped = r[3]; a.AS = r[4]; a.qb = r[5]; a.m_end = r[6]; a.rb = r[7];
      dst[i] = a;
    }
  };
  for (int64_t i = 0; i < n; i++) reads[i].name = trim_readno(names[i]);
  fill(vrec16, vr); fill(jrec16, jr);
  const std::vector<int> vfin = final_sam_order(vr, reads, 0, (size_t)n), jfin = final_sam_order(jr, reads, 0, (size_t)n);
  std::unordered_set<std::string> vnames, share;
  for (int i : vfin) vnames.insert(reads[i].name);
  for (int i : jfin) if (vnames.count(reads[i].name)) share.insert(reads[i].name);
  struct Full { std::string name; int idx; int which; };
  std::vector<Full> full;
  std::vector<int> part[2];
  for (int which = 0; which < 2; which++)
    for (int i : chunk_order(which ? jfin : vfin, nthreads)) {
      if (share.count(reads[i].name)) full.push_back({reads[i].name, i, which});
      else part[which].push_back(i);
    }
  std::stable_sort(full.begin(), full.end(), [](const Full& a, const Full& b) { return a.name < b.name; });
  int64_t t = 0, np = 0;
  for (int i : part[0]) { items[3 * t] = 0; items[3 * t + 1] = i; items[3 * t + 2] = i; t++; }
  for (size_t g = 0; g < full.size();) {
    size_t h = g;
    while (h < full.size() && full[h].name == full[g].name) h++;
    if (h - g == 2) { items[3 * t] = 2; items[3 * t + 1] = full[g].idx; items[3 * t + 2] = full[g + 1].idx; t++; np++; }
    g = h;
  }
  for (int i : part[1]) { items[3 * t] = 1; items[3 * t + 1] = i; items[3 * t + 2] = i; t++; }
  counts[0] = (int64_t)vfin.size(); counts[1] = (int64_t)jfin.size(); counts[2] = (int64_t)part[0].size(); counts[3] = np; counts[4] = (int64_t)part[1].size();
  return t;
}

// Canonical, order-sensitive digest of a list (the twin of catt_result_text_digest in libcatt_b200): per entry FNV-1a over
// seq 0x1f qual 0x1f vs 0x1f js 0x1f cdr3 0x1f able, mixed with the entry's position, summed.  `only_cdr3` skips the reads whose
// cdr3 is "None" (the lists as they stand after catt.jl:581-582); positions then count the kept reads.
uint64_t co_result_digest(void* r, int which, int only_cdr3) {
  RunResult* R = (RunResult*)r;
  uint64_t acc = 0, pos = 0;
  for (const Myread& m : (which ? R->Jpart : R->Vpart)) {
    if (only_cdr3 && m.cdr3 == "None") continue;
    uint64_t h = 0xcbf29ce484222325ull ^ (pos * 0x9E3779B97F4A7C15ull);
    auto mix = [&](const std::string& t) { for (unsigned char b : t) { h ^= b; h *= 0x100000001b3ull; } h ^= 0x1f; h *= 0x100000001b3ull; };
    mix(m.seq); mix(m.qual); mix(m.vs); mix(m.js); mix(m.cdr3);
    h ^= (uint64_t)(m.able ? 1 : 0); h *= 0x100000001b3ull;
    h ^= h >> 29; h *= 0xbf58476d1ce4e5b9ull; h ^= h >> 32;
    acc += h; pos++;
  }
  return acc ^ (pos * 0x94d049bb133111ebull);
}

// per-read alignment records of the run: which=0 V, 1 J; out = n*16 int32 (layout as co_align)
int co_result_records(void* r, int which, int32_t* out) {
  RunResult* R = (RunResult*)r;
  const std::vector<AlnRec>& v = which ? R->jrec : R->vrec;
  for (size_t i = 0; i < v.size(); i++) {
    const AlnRec& a = v[i];
    int32_t* o = out + i * 16;
    o[0] = a.ordinal; o[1] = a.rid; o[2] = a.strand; o[3] = a.mapped; o[4] = a.AS; o[5] = a.qb; o[6] = a.m_end; o[7] = a.rb;
    o[8] = a.qe; o[9] = a.re; o[10] = a.nm; o[11] = a.mi; o[12] = a.ncand; o[13] = a.ntied;
    o[14] = (int32_t)(a.cells & 0xffffffff); o[15] = (int32_t)(a.cells >> 32);
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// nbsorb inner scans.  Parity note: the reference's golden CSV runs nbsorb, but on testSample every
// distinct CDR3 is a root (threshold 1), so linkRLG never executes there and the maybe x highconf scan only ever meets
// an identical sequence: these two restatements follow Jtool.jl line by line and are otherwise "parity unpinned".
// ---------------------------------------------------------------------------------------------------------------
// Jtool.jl:65-74 HammingDistanceG3(s1, s2, upper; skip): true unless more than `upper` of positions skip+1..len-skip differ
static bool hamming_g3(const char* s1, const char* s2, int len, int64_t upper, int skip) {
  int64_t cnt = 0;
  for (int i = skip; i < len - skip; i++)
    if (s1[i] != s2[i] && ++cnt > upper) return false;
  return true;
}

// Jtool.jl:396-421 linkRLG for every leaf (the Threads.@threads loop at Jtool.jl:524-531).  roots / leaves: n x lgt bytes;
// uplimit: n_roots x 7 doubles (Jtool.jl:515-519); statu = number of roots within reach, capped at 2; holder = 1-based index of the
// first such root (1 when there is none).
void co_link_leaves(int lgt, const char* roots, int64_t n_roots, const double* uplimit, const char* leaves, const int64_t* leaf_count,
                    int64_t n_leaves, int32_t* statu, int32_t* holder) {
  #pragma omp parallel for schedule(dynamic, 64)
  for (int64_t l = 0; l < n_leaves; l++) {
    int cnt = 0, hold = 1;
    for (int64_t r = 0; r < n_roots; r++) {
      const double* u = uplimit + r * 7;
      int upper = 1;                                            // argmin(@. abs(uplimit[idy] - leaf[2])): first minimum, 1-based
      double best = std::fabs(u[0] - (double)leaf_count[l]);
      for (int x = 1; x < 7; x++) {
        double dlt = std::fabs(u[x] - (double)leaf_count[l]);
        if (dlt < best) { best = dlt; upper = x + 1; }
      }
      if (hamming_g3(leaves + l * lgt, roots + r * lgt, lgt, upper, 3)) {
        if (++cnt > 1) break;
        hold = (int)(r + 1);
      }
    }
    statu[l] = cnt; holder[l] = hold;
  }
}

// Jtool.jl:553-558: is any target within `upper` mismatches (positions skip+1..len-skip) of the query?
void co_near_any(int lgt, const char* queries, int64_t nq, const char* targets, int64_t nt, int upper, int skip, uint8_t* flag) {
  #pragma omp parallel for schedule(dynamic, 64)
  for (int64_t q = 0; q < nq; q++) {
    uint8_t f = 0;
    for (int64_t t = 0; t < nt && !f; t++) f = hamming_g3(queries + q * lgt, targets + t * lgt, lgt, upper, skip);
    flag[q] = f;
  }
}

}  // extern "C"


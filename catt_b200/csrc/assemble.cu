// assemble.cu — Stage B: from per-read alignment records to Vpart / Jpart (catt.jl:186-339 after the SAM files exist,
// plus the pool / extension / finder!(array) part of catt(), catt.jl:544-577).
//
// Host-side glue restated here (what the reference does with samtools / awk / Julia containers between the stages; none
// of it is a CPU fallback for a kernel):
//   * `samtools sort -t AS | view -F 2308 | tac | awk '!a[$1]++'` (catt.jl:45-48): descending (AS, tid, pos, strand)
//     order, later input first on complete ties, first record per trimmed QNAME                         [SURVEY A7]
//   * get_kmer_fromsam / get_err_fromsam (catt.jl:153-184) as samtools stats prints them                [SURVEY A8]
//   * share_name, the NR % nthreads chunk order, the name-sorted both-mapped pairing (catt.jl:221-319)
//   * Myread string materialisation (catt.jl:104,125-126,302-304; Jtool.jl:364-365; catt.jl:378-379,420-421)
// The per-record work (assignV/J, real_score, finder!, k-mer table, extension) runs in extract.cu / pool.cu.
// Sorting and string work use all host threads (OpenMP); QNAMEs are compared through a 64-bit hash with an exact
// string check on every hash hit.
#include <omp.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <parallel/algorithm>
#include <string_view>

#include "host_types.h"

namespace catt {
namespace {

inline std::string_view trimmed_name(const HostReads& hr, int64_t i) {       // bwa trim_readno
  const char* s = hr.names + hr.name_off[i];
  size_t l = (size_t)(hr.name_off[i + 1] - hr.name_off[i]);
  if (l > 2 && s[l - 2] == '/' && s[l - 1] >= '0' && s[l - 1] <= '9') l -= 2;
  return std::string_view(s, l);
}

inline uint64_t name_hash(std::string_view s) {
  uint64_t h = 0xcbf29ce484222325ull;
  for (unsigned char c : s) { h ^= c; h *= 0x100000001b3ull; }
  h ^= h >> 32; h *= 0xd6e8feb86659fd93ull; h ^= h >> 32;
  return h;
}

// open-addressing set of read indices keyed by trimmed QNAME (hash + exact compare)
struct NameTable {
  std::vector<int32_t> slot;
  uint64_t mask = 0;
  const HostReads* hr = nullptr;
  const uint64_t* nh = nullptr;
  void init(size_t n, const HostReads* h, const uint64_t* hashes) {
    size_t cap = 16;
    while (cap < n * 2 + 2) cap <<= 1;
    slot.assign(cap, -1); mask = cap - 1; hr = h; nh = hashes;
  }
  // index already stored under the name of read i, or -1 (then i is inserted if `insert`)
  int64_t lookup(int64_t i, bool insert) {
    uint64_t s = nh[i] & mask;
    const std::string_view name = trimmed_name(*hr, i);
    while (true) {
      const int32_t v = slot[s];
      if (v < 0) { if (insert) slot[s] = (int32_t)i; return -1; }
      if (nh[v] == nh[i] && trimmed_name(*hr, v) == name) return v;
      s = (s + 1) & mask;
    }
  }
};

struct Keyed { uint64_t key; uint32_t idx; };

// final SAM list of one reference set; `tab` ends up holding the kept names
void final_order(const std::vector<catt_aln>& recs, const HostReads& hr, const uint64_t* nh, std::vector<int64_t>& fin, NameTable& tab) {
  const int64_t n = (int64_t)recs.size();
  std::vector<Keyed> keyed;
  keyed.reserve(n / 2 + 16);
  for (int64_t i = 0; i < n; i++) {
    const catt_aln& a = recs[i];
    if (!a.mapped) continue;
    keyed.push_back(Keyed{((uint64_t)(uint16_t)a.as << 48) | ((uint64_t)(uint16_t)a.rid << 32) | ((uint64_t)(uint16_t)a.rb << 16) | (uint64_t)(a.strand & 1),
                          (uint32_t)i});
  }
  // ascending stable sort followed by `tac` == descending by (key, input index)
  __gnu_parallel::sort(keyed.begin(), keyed.end(), [](const Keyed& x, const Keyed& y) { return x.key != y.key ? x.key > y.key : x.idx > y.idx; });
  tab.init(keyed.size(), &hr, nh);
  fin.clear();
  fin.reserve(keyed.size());
  for (const Keyed& kv : keyed)
    if (tab.lookup(kv.idx, true) < 0) fin.push_back(kv.idx);
}

void chunk_order(const std::vector<int64_t>& fin, int T, std::vector<int64_t>& out) {    // catt.jl:228,256-276
  out.clear();
  if (T <= 1) { out = fin; return; }
  out.reserve(fin.size());
  for (int p = 0; p < T; p++)
    for (size_t nr = (p == 0 ? (size_t)T : (size_t)p); nr <= fin.size(); nr += T) out.push_back(fin[nr - 1]);
}

const char FWD[6] = "ACGTN";
inline char sam_base(const char* s, int len, int i, int strand) {
  if (!strand) return FWD[nt_code(s[i])];
  const uint8_t c = nt_code(s[len - 1 - i]);
  return c > 3 ? 'N' : FWD[3 - c];
}

bool sam_seq_equal(const HostReads& hr, int64_t r1, int s1, int64_t r2, int s2) {
  const int l1 = (int)(hr.seq_off[r1 + 1] - hr.seq_off[r1]), l2 = (int)(hr.seq_off[r2 + 1] - hr.seq_off[r2]);
  if (l1 != l2) return false;
  if (r1 == r2 && s1 == s2) return true;
  const char *a = hr.seqs + hr.seq_off[r1], *b = hr.seqs + hr.seq_off[r2];
  for (int i = 0; i < l1; i++) if (sam_base(a, l1, i, s1) != sam_base(b, l2, i, s2)) return false;
  return true;
}

struct Pending {            // a Myread under construction on the host
  int64_t read; int kind, strand; int seq_off, seq_len, cdr3_off, cdr3_len; int v_id, j_id;
  int q_start1, j_pad;      // J part: qual = qual[1:start-1] * 'G'^j_pad
};

struct SharedPair { std::string_view name; int64_t vi, ji; };

}  // namespace

int stage_b(catt_ctx* c, const HostReads& hr, const DeviceReads& dr, catt_result* R, int* launches) {
  catt_info& info = R->info;
  const std::vector<catt_aln>&vrec = R->rec[0], &jrec = R->rec[1];
  const int64_t n = hr.n;
  const double t0 = now_ms();
  Trace tr;
  std::vector<uint64_t> nh((size_t)n);
  #pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; i++) nh[i] = (vrec[i].mapped || jrec[i].mapped) ? name_hash(trimmed_name(hr, i)) : 0;
  std::vector<int64_t> vfin, jfin;
  NameTable vtab, jtab;
  tr.lap("name hashes");
  final_order(vrec, hr, nh.data(), vfin, vtab);
  final_order(jrec, hr, nh.data(), jfin, jtab);
  tr.lap("final_order x2");
  info.v_final = (int64_t)vfin.size(); info.j_final = (int64_t)jfin.size();
  // get_kmer_fromsam / get_err_fromsam on the final V list, as samtools stats prints them
  int k = c->kmer;
  {
    uint64_t tot = 0, nm = 0, mi = 0;
    #pragma omp parallel for reduction(+ : tot, nm, mi) schedule(static)
    for (int64_t x = 0; x < (int64_t)vfin.size(); x++) {
      const int64_t i = vfin[x];
      tot += (uint64_t)(hr.seq_off[i + 1] - hr.seq_off[i]); nm += (uint64_t)vrec[i].nm; mi += (uint64_t)vrec[i].mi;
    }
    const float avg = vfin.empty() ? 0.f : (float)(tot / (uint64_t)vfin.size());    // samtools: integer division, printed %.0f
    char buf[64];
    snprintf(buf, sizeof buf, "%.0f", avg);
    const double avg_r = atof(buf);
    info.avg_len = avg_r;
    if (k == CATT_KMER_AUTO) {
      if (avg_r >= 50) k = 15;
      if (avg_r >= 75) k = 19;
      if (avg_r >= 100) k = 23;
      if (avg_r >= 150) k = 25;
    }
    const float er = mi ? (float)nm / (float)mi : 0.f;
    snprintf(buf, sizeof buf, "%e", (double)er);
    info.the_err = atof(buf);
    if (!(info.the_err == info.the_err)) info.the_err = 0.0;
  }
  info.kmer = k;
  // share_name = QNAMEs present in both final lists
  std::vector<uint8_t> shared((size_t)n, 0);       // bit 0: V record is shared, bit 1: J record is shared
  std::vector<SharedPair> pairs;
  for (int64_t j : jfin) {
    const int64_t vi = vtab.lookup(j, false);
    if (vi < 0) continue;
    pairs.push_back(SharedPair{trimmed_name(hr, j), vi, j});
    shared[vi] |= 1; shared[j] |= 2;
  }
  info.share = (int64_t)pairs.size();
  tr.lap("stats + share");
  // work items in the reference's processing order
  std::vector<ExtractItem> items;
  std::vector<Pending> pend;
  std::vector<int64_t> order;
  items.reserve(vfin.size() + jfin.size());
  pend.reserve(vfin.size() + jfin.size());
  auto add_part = [&](int which) {
    const std::vector<catt_aln>& recs = which ? jrec : vrec;
    chunk_order(which ? jfin : vfin, c->nthreads, order);
    for (int64_t i : order) {
      if (shared[i] & (which ? 2 : 1)) continue;
      const catt_aln& a = recs[i];
      ExtractItem it{};
      it.read = (uint32_t)i; it.kind = (int16_t)which; it.strand = a.strand; it.qb = a.qb; it.m_end = a.m_end; it.rb = a.rb; it.rid = (int16_t)a.rid;
      items.push_back(it);
      Pending p{};
      p.read = i; p.kind = which; p.strand = a.strand; p.v_id = which ? -1 : a.rid; p.j_id = which ? a.rid : -1;
      p.q_start1 = a.qb + 1; p.j_pad = which ? (c->J.len[a.rid] - (a.rb + 1)) : 0;
      pend.push_back(p);
    }
  };
  add_part(0);
  const size_t n_vpart_items = items.size();
  add_part(1);
  const size_t n_jpart_end = items.size();
  tr.lap("part items");
  // both-mapped reads: sorted by name (V record first, J record second), SEQ equality, clipping V start .. J first-M end
  {
    __gnu_parallel::sort(pairs.begin(), pairs.end(), [](const SharedPair& x, const SharedPair& y) { return x.name < y.name; });
    std::vector<uint8_t> eq(pairs.size());
    #pragma omp parallel for schedule(static)
    for (int64_t x = 0; x < (int64_t)pairs.size(); x++)
      eq[x] = sam_seq_equal(hr, pairs[x].vi, vrec[pairs[x].vi].strand, pairs[x].ji, jrec[pairs[x].ji].strand);
    for (size_t x = 0; x < pairs.size(); x++) {
      if (!eq[x]) { info.pair_seq_mismatch++; continue; }
      const int64_t vi = pairs[x].vi, ji = pairs[x].ji;
      const catt_aln &a = vrec[vi], &b = jrec[ji];
      ExtractItem it{};
      it.read = (uint32_t)vi; it.kind = 2; it.strand = a.strand; it.qb = a.qb; it.m_end = a.m_end; it.rb = a.rb; it.rid = (int16_t)a.rid;
      it.j_m_end = b.m_end; it.j_rid = (int16_t)b.rid;
      items.push_back(it);
      Pending p{};
      p.read = vi; p.kind = 2; p.strand = a.strand; p.v_id = a.rid; p.j_id = b.rid; p.q_start1 = a.qb + 1;
      pend.push_back(p);
    }
  }
  const int64_t n_items = (int64_t)items.size();
  info.ms_host += now_ms() - t0;
  tr.lap("pair sort + items");
  // ---- K5 on the device
  cudaStream_t st = c->stream;
  ExtractItem* d_items = nullptr; ExtractOut* d_eout = nullptr;
  std::vector<ExtractOut> eout(n_items);
  CATT_CUDA(cudaEventRecord(c->ev[5], st));
  if (n_items) {
    int rc0;
    if ((rc0 = c->sb[SB_ITEMS].ensure(n_items * sizeof(ExtractItem))) || (rc0 = c->sb[SB_EOUT].ensure(n_items * sizeof(ExtractOut)))) return rc0;
    d_items = (ExtractItem*)c->sb[SB_ITEMS].p; d_eout = (ExtractOut*)c->sb[SB_EOUT].p;
    CATT_CUDA(cudaMemcpyAsync(d_items, items.data(), n_items * sizeof(ExtractItem), cudaMemcpyHostToDevice, st));
    int rc = launch_extract(c->V.dev, c->J.dev, dr.view(), c->d_finder, d_items, d_eout, n_items, k, c->allow_v, c->allow_j, st, launches);
    if (rc) return rc;
    CATT_CUDA(cudaMemcpyAsync(eout.data(), d_eout, n_items * sizeof(ExtractOut), cudaMemcpyDeviceToHost, st));
  }
  CATT_CUDA(cudaEventRecord(c->ev[6], st));
  CATT_CUDA(cudaStreamSynchronize(st));
  { float ms = 0; cudaEventElapsedTime(&ms, c->ev[5], c->ev[6]); info.ms_extract += ms; }
  tr.lap("extract (alloc+H2D+K5+D2H)");
  // ---- Vpart = kept V part reads, then kept both-mapped reads; Jpart = kept J part reads
  const double t1 = now_ms();
  std::vector<uint32_t> lists[2];
  for (int64_t i = 0; i < (int64_t)n_vpart_items; i++) if (eout[i].status) lists[0].push_back((uint32_t)i);
  for (int64_t i = (int64_t)n_jpart_end; i < n_items; i++) if (eout[i].status) { lists[0].push_back((uint32_t)i); info.full_pushed++; }
  for (int64_t i = (int64_t)n_vpart_items; i < (int64_t)n_jpart_end; i++) if (eout[i].status) lists[1].push_back((uint32_t)i);
  info.vpart = (int64_t)lists[0].size(); info.jpart = (int64_t)lists[1].size();
  const int64_t M = info.vpart + info.jpart;
  // ---- K6/K7: pool + extension of the partial reads
  std::vector<PoolItem> pitems((size_t)M);
  std::vector<int32_t> xi_of((size_t)M, -1);
  std::vector<uint32_t> partial;
  {
    int64_t m = 0;
    for (int side = 0; side < 2; side++)
      for (uint32_t it : lists[side]) {
        Pending& p = pend[it];
        const ExtractOut& e = eout[it];
        p.seq_off = e.seq_off; p.seq_len = e.seq_len; p.cdr3_off = e.cdr3_off; p.cdr3_len = e.cdr3_len;
        PoolItem pi{};
        pi.read = (uint32_t)p.read; pi.strand = (int16_t)p.strand; pi.seq_off = (int16_t)p.seq_off; pi.seq_len = (int16_t)p.seq_len;
        pi.side = (int16_t)side; pi.order = (uint32_t)m; pi.partial = p.cdr3_off < 0;
        if (pi.partial) { xi_of[m] = (int32_t)partial.size(); partial.push_back((uint32_t)m); } else info.direct++;
        pitems[m++] = pi;
      }
  }
  info.left = (int64_t)partial.size();
  info.ms_host += now_ms() - t1;
  tr.lap("lists + pool items");
  std::vector<ExtendOut> xout(partial.size());
  if (M > 0 && k <= 31 && !partial.empty()) {
    uint64_t slots = 1024;
    while (slots < (uint64_t)M * 11ull * 2ull) slots <<= 1;
    KmerTable tab{};
    PoolItem* d_pi = nullptr; uint32_t* d_partial = nullptr; ExtendOut* d_xout = nullptr;
    CATT_CUDA(cudaEventRecord(c->ev[5], st));
    int rc0;
    if ((rc0 = c->sb[SB_KEYS].ensure(slots * sizeof(unsigned long long))) || (rc0 = c->sb[SB_VALS].ensure(slots * sizeof(unsigned long long))) ||
        (rc0 = c->sb[SB_PITEMS].ensure(pitems.size() * sizeof(PoolItem))) || (rc0 = c->sb[SB_PARTIAL].ensure(partial.size() * sizeof(uint32_t))) ||
        (rc0 = c->sb[SB_XOUT].ensure(partial.size() * sizeof(ExtendOut)))) return rc0;
    tab.keys = (unsigned long long*)c->sb[SB_KEYS].p; tab.vals = (unsigned long long*)c->sb[SB_VALS].p;
    d_pi = (PoolItem*)c->sb[SB_PITEMS].p; d_partial = (uint32_t*)c->sb[SB_PARTIAL].p; d_xout = (ExtendOut*)c->sb[SB_XOUT].p;
    tab.mask = slots - 1;
    CATT_CUDA(cudaMemsetAsync(tab.keys, 0xff, slots * sizeof(unsigned long long), st));
    CATT_CUDA(cudaMemsetAsync(tab.vals, 0, slots * sizeof(unsigned long long), st));
    CATT_CUDA(cudaMemcpyAsync(d_pi, pitems.data(), pitems.size() * sizeof(PoolItem), cudaMemcpyHostToDevice, st));
    CATT_CUDA(cudaMemcpyAsync(d_partial, partial.data(), partial.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    int rc = launch_pool(dr.view(), d_pi, M, k, tab, st, launches);
    if (!rc) rc = launch_extend(dr.view(), c->d_finder, d_pi, d_partial, (int64_t)partial.size(), k, tab, d_xout, st, launches);
    if (rc) return rc;
    CATT_CUDA(cudaMemcpyAsync(xout.data(), d_xout, partial.size() * sizeof(ExtendOut), cudaMemcpyDeviceToHost, st));
    CATT_CUDA(cudaEventRecord(c->ev[6], st));
    CATT_CUDA(cudaStreamSynchronize(st));
    { float ms = 0; cudaEventElapsedTime(&ms, c->ev[5], c->ev[6]); info.ms_pool += ms; }
  } else {
    for (auto& x : xout) { x.n_ext = 0; x.cdr3_off = -1; x.cdr3_len = 0; x.able = 1; }
  }
  tr.lap("pool + extend (alloc..D2H)");
  // ---- materialise Myread strings: sizes, prefix sums, parallel fill into two flat arenas
  const double t2 = now_ms();
  std::vector<uint32_t> flat((size_t)M);
  { int64_t m = 0; for (int side = 0; side < 2; side++) for (uint32_t it : lists[side]) flat[m++] = it; }
  std::vector<int64_t> soff((size_t)M + 1, 0), qoff((size_t)M + 1, 0);
  int64_t rescued = 0;
  #pragma omp parallel for schedule(static) reduction(+ : rescued)
  for (int64_t m = 0; m < M; m++) {
    const Pending& p = pend[flat[m]];
    int n_ext = 0, c_off = p.cdr3_off, c_len = p.cdr3_len;
    if (xi_of[m] >= 0) { const ExtendOut& x = xout[xi_of[m]]; n_ext = x.n_ext; c_off = x.cdr3_off; c_len = x.cdr3_len; rescued += c_off >= 0; }
    const int base_q = p.kind == 1 ? (p.q_start1 - 1 + std::max(0, p.j_pad)) : p.seq_len;
    soff[m + 1] = p.seq_len + n_ext + 1;
    qoff[m + 1] = (c_off >= 0 ? c_len : base_q + n_ext) + 1;
  }
  info.rescued += rescued;
  for (int64_t m = 0; m < M; m++) { soff[m + 1] += soff[m]; qoff[m + 1] += qoff[m]; }
  R->seq_arena.resize((size_t)soff[M] + 1);
  R->qual_arena.resize((size_t)qoff[M] + 1);
  R->part[0].resize((size_t)info.vpart);
  R->part[1].resize((size_t)info.jpart);
  static const char NT[] = "ACGT";
  #pragma omp parallel for schedule(static)
  for (int64_t m = 0; m < M; m++) {
    const Pending& p = pend[flat[m]];
    const int side = m < info.vpart ? 0 : 1;
    const int len = (int)(hr.seq_off[p.read + 1] - hr.seq_off[p.read]);
    const char* rs = hr.seqs + hr.seq_off[p.read];
    const char* rq = hr.quals ? hr.quals + hr.seq_off[p.read] : nullptr;
    int n_ext = 0, c_off = p.cdr3_off, c_len = p.cdr3_len;
    const ExtendOut* x = nullptr;
    if (xi_of[m] >= 0) { x = &xout[xi_of[m]]; n_ext = x->n_ext; c_off = x->cdr3_off; c_len = x->cdr3_len; }
    const int lead = side == 1 ? n_ext : 0;                       // a left extension is prepended
    char* s = &R->seq_arena[(size_t)soff[m]];
    for (int i = 0; i < lead; i++) s[i] = NT[x->ext[n_ext - 1 - i] & 3];          // reverse(ext) * seq
    for (int i = 0; i < p.seq_len; i++) s[lead + i] = sam_base(rs, len, p.seq_off + i, p.strand);
    if (side == 0) for (int i = 0; i < n_ext; i++) s[p.seq_len + i] = NT[x->ext[i] & 3];
    const int slen = p.seq_len + n_ext;
    s[slen] = 0;
    // quality: the Myread's string before finder! slices it
    const int base_q = p.kind == 1 ? (p.q_start1 - 1 + std::max(0, p.j_pad)) : p.seq_len;
    const int qfull = base_q + n_ext;
    auto sq = [&](int i) { return rq ? (p.strand ? rq[len - 1 - i] : rq[i]) : 'I'; };
    auto qat = [&](int i) -> char {                               // i in [0, qfull)
      int b = i - lead;
      if (b < 0 || b >= base_q) return 'G';                      // extension padding
      if (p.kind == 1) return b < p.q_start1 - 1 ? sq(b) : 'G';  // catt.jl:126
      return sq(p.seq_off + b);
    };
    char* q = &R->qual_arena[(size_t)qoff[m]];
    uint8_t flags = (uint8_t)(p.kind == 0 ? 1 : p.kind == 1 ? 2 : 4);
    int qlen;
    if (c_off >= 0) {                                             // rd.qual = rd.qual[range] (Jtool.jl:365,383)
      qlen = c_len;
      for (int i = 0; i < c_len; i++) {
        if (c_off + i < qfull) q[i] = qat(c_off + i); else { q[i] = 'G'; flags |= 8; }    // reference: BoundsError
      }
    } else {
      qlen = qfull;
      for (int i = 0; i < qfull; i++) q[i] = qat(i);
    }
    q[qlen] = 0;
    catt_read rd{};
    rd.seq = s; rd.qual = q; rd.seq_len = slen; rd.qual_len = qlen;
    rd.v_id = p.v_id; rd.j_id = p.j_id; rd.cdr3_off = c_off; rd.cdr3_len = c_off >= 0 ? c_len : 0;
    rd.ordinal = (int32_t)p.read; rd.able = 1; rd.flags = flags;
    R->part[side][side == 0 ? m : m - info.vpart] = rd;
  }
  info.ms_host += now_ms() - t2;
  tr.lap("materialise");
  return CATT_OK;
}

}  // namespace catt

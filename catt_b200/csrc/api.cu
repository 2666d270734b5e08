// api.cu — the C ABI of libcatt_b200 (include/catt_b200.h): context, Stage A (map2align on the GPU), Stage B
// (read_alignrs + pool/extension) and result assembly.
//
// Host-side logic restated here (none of it is a CPU fallback for a kernel — it is the glue the reference does
// with samtools/awk/Julia containers between the stages):
//   * `samtools sort -t AS | view -F 2308 | tac | awk '!a[$1]++'` (catt.jl:45-48): descending (AS, tid, pos, strand)
//     order, later input first on complete ties, first record per trimmed QNAME            [SURVEY A7]
//   * get_kmer_fromsam / get_err_fromsam (catt.jl:153-184) as samtools stats prints them   [SURVEY A8]
//   * share_name, the NR % nthreads chunk order, the name-sorted both-mapped pairing (catt.jl:221-319)
//   * Myread string materialisation (catt.jl:104,125-126,302-304; Jtool.jl:364-365; catt.jl:378-379,420-421).
#include <stdio.h>
#include <string.h>
#include <zlib.h>

#include <algorithm>
#include <chrono>
#include <memory>
#include <string>
#include <string_view>
#include <unordered_map>
#include <vector>

#include "internal.cuh"

namespace catt {
const char* last_error();
bool slurp_file(const std::string& path, std::string& out);
}  // namespace catt

using namespace catt;

struct DeviceReads {
  uint32_t* words = nullptr; uint32_t* off = nullptr; int32_t* len = nullptr;
  int64_t n = 0; int max_len = 0;
  ReadsDev view() const { return ReadsDev{words, off, len, n, max_len}; }
  void release() { cudaFree(words); cudaFree(off); cudaFree(len); words = nullptr; off = nullptr; len = nullptr; }
};

struct catt_ctx {
  RefSet V, J;
  FinderDev finder{};
  FinderDev* d_finder = nullptr;
  int min_score = 12, kmer = CATT_KMER_AUTO, nthreads = 1, allow_v = 3, allow_j = 9, device = 0, sm_count = 148;
  cudaStream_t stream = nullptr;
  std::vector<std::string> d_names, d_seqs;
  char* d_dseq = nullptr; int32_t* d_doff = nullptr;
  AlignBuffers ab[2];
  cudaEvent_t ev[8]{};
};

struct catt_batch {
  DeviceReads dr;
  int64_t first_ordinal = 0;
  int64_t ntasks[2] = {0, 0};
  catt_info info{};
};

struct catt_result {
  catt_info info{};
  std::vector<catt_read> part[2];
  std::vector<catt_aln> rec[2];
  std::vector<std::unique_ptr<std::string>> arena;
};

namespace {

double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int upload_reads(const PackedReads& pr, int64_t n, DeviceReads* dr, cudaStream_t st) {
  dr->n = n; dr->max_len = pr.max_len;
  CATT_CUDA(cudaMalloc(&dr->words, std::max<size_t>(pr.words.size(), 4) * sizeof(uint32_t)));
  CATT_CUDA(cudaMalloc(&dr->off, (n + 1) * sizeof(uint32_t)));
  CATT_CUDA(cudaMalloc(&dr->len, std::max<int64_t>(n, 1) * sizeof(int32_t)));
  CATT_CUDA(cudaMemcpyAsync(dr->words, pr.words.data(), pr.words.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
  CATT_CUDA(cudaMemcpyAsync(dr->off, pr.off.data(), (n + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
  CATT_CUDA(cudaMemcpyAsync(dr->len, pr.len.data(), n * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  return CATT_OK;
}

int ensure_align_buffers(AlignBuffers& ab, const RefDev& ref, int64_t n) {
  if (n <= ab.cap_reads) return CATT_OK;
  cudaFree(ab.candbits); cudaFree(ab.ntask); cudaFree(ab.task_off); cudaFree(ab.recs); cudaFree(ab.gapped_list);
  ab.cap_reads = n + n / 8 + 64;
  CATT_CUDA(cudaMalloc(&ab.candbits, ab.cap_reads * ref.CW * sizeof(uint32_t)));
  CATT_CUDA(cudaMalloc(&ab.ntask, (ab.cap_reads + 1) * sizeof(uint32_t)));
  CATT_CUDA(cudaMalloc(&ab.task_off, (ab.cap_reads + 1) * sizeof(uint32_t)));
  CATT_CUDA(cudaMalloc(&ab.recs, ab.cap_reads * sizeof(catt_aln)));
  CATT_CUDA(cudaMalloc(&ab.gapped_list, ab.cap_reads * sizeof(uint32_t)));
  if (!ab.counters) CATT_CUDA(cudaMalloc(&ab.counters, 8 * sizeof(unsigned long long)));
  return CATT_OK;
}

void free_align_buffers(AlignBuffers& ab) {
  cudaFree(ab.candbits); cudaFree(ab.ntask); cudaFree(ab.task_off); cudaFree(ab.tasks); cudaFree(ab.task_res); cudaFree(ab.recs);
  cudaFree(ab.gapped_list); cudaFree(ab.counters); cudaFree(ab.cub_tmp); cudaFree(ab.seed_scratch);
  ab = AlignBuffers{};
}

// Stage A for one reference set; records stay in ab.recs.  stage_ms[0..2] = seed, sw, select+traceback (device, ms).
int align_refset(catt_ctx* c, int which, const ReadsDev& reads, int64_t first_ordinal, catt_info* info, int* launches) {
  RefSet& rs = which ? c->J : c->V;
  AlignBuffers& ab = c->ab[which];
  cudaStream_t st = c->stream;
  int rc;
  if ((rc = ensure_align_buffers(ab, rs.dev, reads.n))) return rc;
  CATT_CUDA(cudaMemsetAsync(ab.counters, 0, 8 * sizeof(unsigned long long), st));
  CATT_CUDA(cudaMemsetAsync(ab.ntask, 0, (reads.n + 1) * sizeof(uint32_t), st));
  CATT_CUDA(cudaEventRecord(c->ev[0], st));
  if ((rc = launch_seed(rs.dev, reads, ab, c->sm_count, st, launches))) return rc;
  CATT_CUDA(cudaEventRecord(c->ev[1], st));
  int64_t ntasks = 0;
  if ((rc = launch_expand(rs.dev, reads, ab, &ntasks, st, launches))) return rc;
  CATT_CUDA(cudaEventRecord(c->ev[2], st));
  if ((rc = launch_sw(rs.dev, reads, ab, ntasks, c->sm_count, st, launches))) return rc;
  CATT_CUDA(cudaEventRecord(c->ev[3], st));
  if ((rc = launch_select(rs.dev, reads, ab, first_ordinal, c->min_score, st, launches))) return rc;
  unsigned long long cnt[8];
  CATT_CUDA(cudaMemcpyAsync(cnt, ab.counters, sizeof cnt, cudaMemcpyDeviceToHost, st));
  CATT_CUDA(cudaStreamSynchronize(st));
  if (cnt[6] > 0) { set_error("%llu reads exceed the seeding run-list capacity", cnt[6]); return CATT_E_LIMIT; }
  if ((rc = launch_traceback(rs.dev, reads, ab, (int64_t)cnt[2], st, launches))) return rc;
  CATT_CUDA(cudaEventRecord(c->ev[4], st));
  CATT_CUDA(cudaEventSynchronize(c->ev[4]));
  float ms_seed = 0, ms_exp = 0, ms_sw = 0, ms_sel = 0;
  cudaEventElapsedTime(&ms_seed, c->ev[0], c->ev[1]);
  cudaEventElapsedTime(&ms_exp, c->ev[1], c->ev[2]);
  cudaEventElapsedTime(&ms_sw, c->ev[2], c->ev[3]);
  cudaEventElapsedTime(&ms_sel, c->ev[3], c->ev[4]);
  if (info) {
    info->ms_seed += ms_seed + ms_exp; info->ms_sw += ms_sw; info->ms_select += ms_sel;
    if (which) { info->cand_j = (int64_t)cnt[0]; info->cells_j = (int64_t)cnt[1]; info->gapped_j = (int64_t)cnt[2]; info->j_hits = (int64_t)cnt[3]; }
    else { info->cand_v = (int64_t)cnt[0]; info->cells_v = (int64_t)cnt[1]; info->gapped_v = (int64_t)cnt[2]; info->v_hits = (int64_t)cnt[3]; }
  }
  return CATT_OK;
}

// ---- FASTQ / FASTA input (bwa's kseq: name = header up to whitespace) ---------------------------------------------
int parse_fastx(const std::string& path, HostReads* hr) {
  std::string txt;
  if (!slurp_file(path, txt)) { set_error("cannot read %s", path.c_str()); return CATT_E_IO; }
  size_t p = 0;
  const size_t N = txt.size();
  auto line = [&](size_t& b, size_t& e) {
    if (p >= N) return false;
    const char* nl = (const char*)memchr(txt.data() + p, '\n', N - p);
    b = p; e = nl ? (size_t)(nl - txt.data()) : N;
    p = e + 1;
    if (e > b && txt[e - 1] == '\r') e--;
    return true;
  };
  hr->own_name_off.push_back(0); hr->own_seq_off.push_back(0);
  size_t b, e;
  while (line(b, e)) {
    if (e == b) continue;
    const char tag = txt[b];
    if (tag != '@' && tag != '>') { set_error("%s: malformed FASTA/FASTQ record near byte %zu", path.c_str(), b); return CATT_E_IO; }
    size_t k = b + 1;
    while (k < e && txt[k] != ' ' && txt[k] != '\t') k++;
    hr->own_names.append(txt, b + 1, k - b - 1);
    hr->own_name_off.push_back((int64_t)hr->own_names.size());
    size_t sb, se;
    if (!line(sb, se)) { set_error("%s: truncated record", path.c_str()); return CATT_E_IO; }
    hr->own_seqs.append(txt, sb, se - sb);
    hr->own_seq_off.push_back((int64_t)hr->own_seqs.size());
    if (tag == '@') {
      size_t pb, pe, qb, qe;
      if (!line(pb, pe) || !line(qb, qe) || qe - qb != se - sb) { set_error("%s: truncated or inconsistent FASTQ record", path.c_str()); return CATT_E_IO; }
      hr->own_quals.append(txt, qb, qe - qb);
    } else {
      hr->own_quals.append(se - sb, 'I');
    }
  }
  hr->n = (int64_t)hr->own_seq_off.size() - 1;
  hr->names = hr->own_names.data(); hr->name_off = hr->own_name_off.data();
  hr->seqs = hr->own_seqs.data(); hr->seq_off = hr->own_seq_off.data();
  hr->quals = hr->own_quals.data();
  return CATT_OK;
}

inline std::string_view trimmed_name(const HostReads& hr, int64_t i) {       // bwa trim_readno
  const char* s = hr.names + hr.name_off[i];
  size_t l = (size_t)(hr.name_off[i + 1] - hr.name_off[i]);
  if (l > 2 && s[l - 2] == '/' && s[l - 1] >= '0' && s[l - 1] <= '9') l -= 2;
  return std::string_view(s, l);
}

inline char sam_base(const HostReads& hr, int64_t r, int len, int i, int strand) {
  const char* s = hr.seqs + hr.seq_off[r];
  static const char NT[] = "ACGTN";
  if (!strand) return NT[nt_code(s[i])];
  const uint8_t c = nt_code(s[len - 1 - i]);
  return c > 3 ? 'N' : NT[3 - c];
}

bool sam_seq_equal(const HostReads& hr, int64_t r1, int s1, int64_t r2, int s2) {
  const int l1 = (int)(hr.seq_off[r1 + 1] - hr.seq_off[r1]), l2 = (int)(hr.seq_off[r2 + 1] - hr.seq_off[r2]);
  if (l1 != l2) return false;
  if (r1 == r2 && s1 == s2) return true;
  for (int i = 0; i < l1; i++) if (sam_base(hr, r1, l1, i, s1) != sam_base(hr, r2, l2, i, s2)) return false;
  return true;
}

// final SAM list of one reference set
void final_order(const std::vector<catt_aln>& recs, const HostReads& hr, std::vector<int64_t>& fin) {
  std::vector<std::pair<uint64_t, int64_t>> keyed;
  for (int64_t i = 0; i < (int64_t)recs.size(); i++) {
    const catt_aln& a = recs[i];
    if (!a.mapped) continue;
    const uint64_t k = ((uint64_t)(uint16_t)a.as << 48) | ((uint64_t)(uint16_t)a.rid << 32) | ((uint64_t)(uint16_t)a.rb << 16) | (uint64_t)(a.strand & 1);
    keyed.emplace_back(k, i);
  }
  // ascending stable sort then reverse == descending by (key, input index)
  std::sort(keyed.begin(), keyed.end(), [](const std::pair<uint64_t, int64_t>& x, const std::pair<uint64_t, int64_t>& y) {
    return x.first != y.first ? x.first > y.first : x.second > y.second;
  });
  std::unordered_map<std::string_view, char> seen;
  seen.reserve(keyed.size() * 2);
  fin.clear();
  for (auto& kv : keyed)
    if (seen.emplace(trimmed_name(hr, kv.second), 1).second) fin.push_back(kv.second);
}

void chunk_order(const std::vector<int64_t>& fin, int T, std::vector<int64_t>& out) {    // catt.jl:228,256-276
  out.clear();
  if (T <= 1) { out = fin; return; }
  for (int p = 0; p < T; p++)
    for (size_t nr = 1; nr <= fin.size(); nr++)
      if ((int)(nr % T) == p) out.push_back(fin[nr - 1]);
}

const char* arena_str(catt_result* R, std::string&& s) {
  R->arena.emplace_back(new std::string(std::move(s)));
  return R->arena.back()->c_str();
}

struct Pending {            // a Myread under construction on the host
  int64_t read; int kind, strand; int seq_off, seq_len, cdr3_off, cdr3_len; int v_id, j_id;
  int q_start1, j_pad;      // J part: qual = qual[1:start-1] * 'G'^j_pad
  bool able;
};

int stage_b(catt_ctx* c, const HostReads& hr, const DeviceReads& dr, catt_result* R, int* launches) {
  catt_info& info = R->info;
  const std::vector<catt_aln>&vrec = R->rec[0], &jrec = R->rec[1];
  const double t0 = now_ms();
  std::vector<int64_t> vfin, jfin;
  final_order(vrec, hr, vfin);
  final_order(jrec, hr, jfin);
  info.v_final = (int64_t)vfin.size(); info.j_final = (int64_t)jfin.size();
  // get_kmer_fromsam / get_err_fromsam on the final V list, as samtools stats prints them
  int k = c->kmer;
  {
    uint64_t tot = 0, nm = 0, mi = 0;
    for (int64_t i : vfin) { tot += (uint64_t)(hr.seq_off[i + 1] - hr.seq_off[i]); nm += (uint64_t)vrec[i].nm; mi += (uint64_t)vrec[i].mi; }
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
  std::unordered_map<std::string_view, int64_t> vname;
  vname.reserve(vfin.size() * 2);
  for (int64_t i : vfin) vname.emplace(trimmed_name(hr, i), i);
  std::unordered_map<std::string_view, std::pair<int64_t, int64_t>> share;   // name -> (V read, J read)
  for (int64_t j : jfin) {
    auto it = vname.find(trimmed_name(hr, j));
    if (it != vname.end()) share.emplace(it->first, std::make_pair(it->second, j));
  }
  info.share = (int64_t)share.size();
  // work items in the reference's processing order
  std::vector<ExtractItem> items;
  std::vector<Pending> pend;
  std::vector<int64_t> order;
  auto add_part = [&](int which) {
    const std::vector<catt_aln>& recs = which ? jrec : vrec;
    chunk_order(which ? jfin : vfin, c->nthreads, order);
    for (int64_t i : order) {
      if (share.count(trimmed_name(hr, i))) continue;
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
  // both-mapped reads: sorted by name (stable: V record then J record), SEQ equality, clipping V start .. J first-M end
  {
    std::vector<std::pair<std::string_view, std::pair<int64_t, int64_t>>> pairs(share.begin(), share.end());
    std::sort(pairs.begin(), pairs.end(), [](const auto& x, const auto& y) { return x.first < y.first; });
    for (auto& pr : pairs) {
      const int64_t vi = pr.second.first, ji = pr.second.second;
      const catt_aln &a = vrec[vi], &b = jrec[ji];
      if (!sam_seq_equal(hr, vi, a.strand, ji, b.strand)) { info.pair_seq_mismatch++; continue; }
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
  // ---- K5 on the device
  cudaStream_t st = c->stream;
  ExtractItem* d_items = nullptr; ExtractOut* d_eout = nullptr;
  std::vector<ExtractOut> eout(n_items);
  CATT_CUDA(cudaEventRecord(c->ev[5], st));
  if (n_items) {
    CATT_CUDA(cudaMalloc(&d_items, n_items * sizeof(ExtractItem)));
    CATT_CUDA(cudaMalloc(&d_eout, n_items * sizeof(ExtractOut)));
    CATT_CUDA(cudaMemcpyAsync(d_items, items.data(), n_items * sizeof(ExtractItem), cudaMemcpyHostToDevice, st));
    int rc = launch_extract(c->V.dev, c->J.dev, dr.view(), c->d_finder, d_items, d_eout, n_items, k, c->allow_v, c->allow_j, st, launches);
    if (rc) return rc;
    CATT_CUDA(cudaMemcpyAsync(eout.data(), d_eout, n_items * sizeof(ExtractOut), cudaMemcpyDeviceToHost, st));
  }
  CATT_CUDA(cudaEventRecord(c->ev[6], st));
  CATT_CUDA(cudaStreamSynchronize(st));
  cudaFree(d_items); cudaFree(d_eout);
  { float ms = 0; cudaEventElapsedTime(&ms, c->ev[5], c->ev[6]); info.ms_extract += ms; }
  // ---- Vpart = kept V part reads, then kept both-mapped reads; Jpart = kept J part reads
  const double t1 = now_ms();
  std::vector<Pending> lists[2];
  for (int64_t i = 0; i < n_items; i++) {
    if (!eout[i].status) continue;
    Pending p = pend[i];
    p.seq_off = eout[i].seq_off; p.seq_len = eout[i].seq_len; p.cdr3_off = eout[i].cdr3_off; p.cdr3_len = eout[i].cdr3_len; p.able = eout[i].able != 0;
    const bool is_j = (size_t)i >= n_vpart_items && (size_t)i < n_jpart_end;
    if (!is_j && (size_t)i < n_vpart_items) lists[0].push_back(p);
    else if (is_j) lists[1].push_back(p);
  }
  for (int64_t i = (int64_t)n_jpart_end; i < n_items; i++) {
    if (!eout[i].status) continue;
    Pending p = pend[i];
    p.seq_off = eout[i].seq_off; p.seq_len = eout[i].seq_len; p.cdr3_off = eout[i].cdr3_off; p.cdr3_len = eout[i].cdr3_len; p.able = eout[i].able != 0;
    lists[0].push_back(p);
    info.full_pushed++;
  }
  info.vpart = (int64_t)lists[0].size(); info.jpart = (int64_t)lists[1].size();
  // ---- K6/K7: pool + extension of the partial reads
  std::vector<PoolItem> pitems;
  std::vector<uint32_t> partial;
  for (int side = 0; side < 2; side++)
    for (const Pending& p : lists[side]) {
      PoolItem pi{};
      pi.read = (uint32_t)p.read; pi.strand = (int16_t)p.strand; pi.seq_off = (int16_t)p.seq_off; pi.seq_len = (int16_t)p.seq_len;
      pi.side = (int16_t)side; pi.order = (uint32_t)pitems.size(); pi.partial = p.cdr3_off < 0;
      if (pi.partial) partial.push_back((uint32_t)pitems.size()); else info.direct++;
      pitems.push_back(pi);
    }
  info.left = (int64_t)partial.size();
  info.ms_host += now_ms() - t1;
  std::vector<ExtendOut> xout(partial.size());
  if (!pitems.empty() && k <= 31 && !partial.empty()) {
    uint64_t slots = 1024;
    while (slots < pitems.size() * 11ull * 2ull) slots <<= 1;
    KmerTable tab{};
    PoolItem* d_pi = nullptr; uint32_t* d_partial = nullptr; ExtendOut* d_xout = nullptr;
    CATT_CUDA(cudaEventRecord(c->ev[5], st));
    CATT_CUDA(cudaMalloc(&tab.keys, slots * sizeof(unsigned long long)));
    CATT_CUDA(cudaMalloc(&tab.vals, slots * sizeof(unsigned long long)));
    tab.mask = slots - 1;
    CATT_CUDA(cudaMemsetAsync(tab.keys, 0xff, slots * sizeof(unsigned long long), st));
    CATT_CUDA(cudaMemsetAsync(tab.vals, 0, slots * sizeof(unsigned long long), st));
    CATT_CUDA(cudaMalloc(&d_pi, pitems.size() * sizeof(PoolItem)));
    CATT_CUDA(cudaMalloc(&d_partial, partial.size() * sizeof(uint32_t)));
    CATT_CUDA(cudaMalloc(&d_xout, partial.size() * sizeof(ExtendOut)));
    CATT_CUDA(cudaMemcpyAsync(d_pi, pitems.data(), pitems.size() * sizeof(PoolItem), cudaMemcpyHostToDevice, st));
    CATT_CUDA(cudaMemcpyAsync(d_partial, partial.data(), partial.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    int rc = launch_pool(dr.view(), d_pi, (int64_t)pitems.size(), k, tab, st, launches);
    if (!rc) rc = launch_extend(dr.view(), c->d_finder, d_pi, d_partial, (int64_t)partial.size(), k, tab, d_xout, st, launches);
    if (rc) return rc;
    CATT_CUDA(cudaMemcpyAsync(xout.data(), d_xout, partial.size() * sizeof(ExtendOut), cudaMemcpyDeviceToHost, st));
    CATT_CUDA(cudaEventRecord(c->ev[6], st));
    CATT_CUDA(cudaStreamSynchronize(st));
    { float ms = 0; cudaEventElapsedTime(&ms, c->ev[5], c->ev[6]); info.ms_pool += ms; }
    cudaFree(tab.keys); cudaFree(tab.vals); cudaFree(d_pi); cudaFree(d_partial); cudaFree(d_xout);
  } else {
    for (auto& x : xout) { x.n_ext = 0; x.cdr3_off = -1; x.cdr3_len = 0; x.able = 1; }
  }
  // ---- materialise Myread strings
  const double t2 = now_ms();
  static const char NT[] = "ACGT";
  size_t pidx = 0, xi = 0;
  for (int side = 0; side < 2; side++) {
    R->part[side].reserve(lists[side].size());
    for (const Pending& p : lists[side]) {
      const PoolItem& pi = pitems[pidx++];
      const int len = (int)(hr.seq_off[p.read + 1] - hr.seq_off[p.read]);
      std::string seq((size_t)p.seq_len, 'N');
      for (int i = 0; i < p.seq_len; i++) seq[i] = sam_base(hr, p.read, len, p.seq_off + i, p.strand);
      // quality in SAM orientation
      const char* q = hr.quals ? hr.quals + hr.seq_off[p.read] : nullptr;
      auto sq = [&](int i) { return q ? (p.strand ? q[len - 1 - i] : q[i]) : 'I'; };
      std::string qual;
      if (p.kind == 1) {                                   // catt.jl:126
        for (int i = 0; i < p.q_start1 - 1; i++) qual.push_back(sq(i));
        qual.append((size_t)std::max(0, p.j_pad), 'G');
      } else {
        for (int i = 0; i < p.seq_len; i++) qual.push_back(sq(p.seq_off + i));
      }
      int c_off = p.cdr3_off, c_len = p.cdr3_len;
      if (pi.partial) {
        const ExtendOut& x = xout[xi++];
        if (x.n_ext > 0) {
          std::string add((size_t)x.n_ext, 'A');
          for (int i = 0; i < x.n_ext; i++) add[i] = NT[x.ext[i] & 3];
          if (side == 0) { seq += add; qual.append((size_t)x.n_ext, 'G'); }
          else { std::reverse(add.begin(), add.end()); seq = add + seq; qual = std::string((size_t)x.n_ext, 'G') + qual; }
        }
        c_off = x.cdr3_off; c_len = x.cdr3_len;
        if (c_off >= 0) info.rescued++;
      }
      uint8_t flags = (uint8_t)(p.kind == 0 ? 1 : p.kind == 1 ? 2 : 4);
      if (c_off >= 0) {                                    // rd.qual = rd.qual[range] (Jtool.jl:365,383)
        std::string qn;
        for (int i = c_off; i < c_off + c_len; i++) {
          if (i < (int)qual.size()) qn.push_back(qual[i]); else { qn.push_back('G'); flags |= 8; }   // reference: BoundsError
        }
        qual.swap(qn);
      }
      catt_read rd{};
      rd.seq_len = (int32_t)seq.size(); rd.qual_len = (int32_t)qual.size();
      rd.seq = arena_str(R, std::move(seq)); rd.qual = arena_str(R, std::move(qual));
      rd.v_id = p.v_id; rd.j_id = p.j_id; rd.cdr3_off = c_off; rd.cdr3_len = c_off >= 0 ? c_len : 0;
      rd.ordinal = (int32_t)p.read; rd.able = 1; rd.flags = flags;
      R->part[side].push_back(rd);
    }
  }
  info.ms_host += now_ms() - t2;
  return CATT_OK;
}

int run_all(catt_ctx* c, const HostReads& hr, catt_result** out) {
  std::unique_ptr<catt_result> R(new catt_result());
  int launches = 0;
  const double t0 = now_ms();
  PackedReads pr;
  int rc = pack_reads(hr.n, hr.seqs, hr.seq_off, &pr);
  if (rc) return rc;
  R->info.ms_host += now_ms() - t0;
  DeviceReads dr;
  CATT_CUDA(cudaEventRecord(c->ev[5], c->stream));
  if ((rc = upload_reads(pr, hr.n, &dr, c->stream))) return rc;
  CATT_CUDA(cudaEventRecord(c->ev[6], c->stream));
  CATT_CUDA(cudaStreamSynchronize(c->stream));
  { float ms = 0; cudaEventElapsedTime(&ms, c->ev[5], c->ev[6]); R->info.ms_h2d += ms; }
  R->info.n_reads = hr.n;
  for (int which = 0; which < 2 && rc == CATT_OK; which++) {
    rc = align_refset(c, which, dr.view(), 0, &R->info, &launches);
    if (rc) break;
    R->rec[which].resize(hr.n);
    const double td = now_ms();
    cudaError_t e = cudaMemcpy(R->rec[which].data(), c->ab[which].recs, hr.n * sizeof(catt_aln), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = cuda_fail(e, "records D2H", __FILE__, __LINE__);
    R->info.ms_d2h += now_ms() - td;
  }
  if (rc == CATT_OK) rc = stage_b(c, hr, dr, R.get(), &launches);
  dr.release();
  if (rc) return rc;
  R->info.kernel_launches = launches;
  *out = R.release();
  return CATT_OK;
}

std::string dup(const char* s) { return s ? std::string(s) : std::string(); }

}  // namespace

// ================================================================================================================
// C ABI
// ================================================================================================================
extern "C" {

const char* catt_last_error(void) { return catt::last_error(); }
const char* catt_version(void) { return "catt_b200 0.1 (sm_100a)"; }

int catt_ctx_create(catt_ctx** out, const catt_config* cfg) {
  if (!out || !cfg || !cfg->v_fasta || !cfg->j_fasta || !cfg->cmotif || !cfg->fmotif || !cfg->inner_c || !cfg->inner_f) {
    set_error("catt_ctx_create: null argument");
    return CATT_E_ARG;
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    set_error("no usable CUDA device (%s); libcatt_b200 has no CPU fallback", e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    return CATT_E_CUDA;
  }
  if (cfg->device < 0 || cfg->device >= ndev) { set_error("device %d out of range (have %d)", cfg->device, ndev); return CATT_E_ARG; }
  CATT_CUDA(cudaSetDevice(cfg->device));
  std::unique_ptr<catt_ctx> c(new catt_ctx());
  c->device = cfg->device;
  cudaDeviceProp prop;
  CATT_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
  c->sm_count = prop.multiProcessorCount;
  c->min_score = cfg->min_score; c->kmer = cfg->kmer; c->nthreads = cfg->julia_nthreads < 1 ? 1 : cfg->julia_nthreads;
  c->allow_v = cfg->allow_v; c->allow_j = cfg->allow_j;
  int rc;
  if ((rc = c->V.load(cfg->v_fasta)) || (rc = c->J.load(cfg->j_fasta))) return rc;
  if ((rc = c->V.upload()) || (rc = c->J.upload())) return rc;
  const char* res[M_COUNT] = {cfg->cmotif, cfg->fmotif, cfg->inner_c, cfg->inner_f, "(CAS|CSA|CAW|CAT|CSV|CAI|CAR|CAG|CSG)",
                              "(QYF|QFF|AFF|LFF|YTF|QHF|LHF|IYF|LTF)"};     // Jtool.jl:288,294
  for (int i = 0; i < M_COUNT; i++) if ((rc = compile_motif(res[i], &c->finder.m[i]))) return rc;
  c->finder.coffset = cfg->coffset; c->finder.foffset = cfg->foffset;
  CATT_CUDA(cudaMalloc(&c->d_finder, sizeof(FinderDev)));
  CATT_CUDA(cudaMemcpy(c->d_finder, &c->finder, sizeof(FinderDev), cudaMemcpyHostToDevice));
  if (cfg->n_d > 0) {
    std::string cat; std::vector<int32_t> off{0};
    for (int i = 0; i < cfg->n_d; i++) {
      c->d_names.push_back(dup(cfg->d_names[i])); c->d_seqs.push_back(dup(cfg->d_seqs[i]));
      if (c->d_seqs.back().size() > 32) { set_error("D segment %s longer than 32 nt", c->d_names.back().c_str()); return CATT_E_LIMIT; }
      cat += c->d_seqs.back(); off.push_back((int32_t)cat.size());
    }
    CATT_CUDA(cudaMalloc(&c->d_dseq, cat.size() + 16));
    CATT_CUDA(cudaMalloc(&c->d_doff, off.size() * sizeof(int32_t)));
    CATT_CUDA(cudaMemcpy(c->d_dseq, cat.data(), cat.size(), cudaMemcpyHostToDevice));
    CATT_CUDA(cudaMemcpy(c->d_doff, off.data(), off.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  }
  CATT_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  for (auto& ev : c->ev) CATT_CUDA(cudaEventCreate(&ev));
  *out = c.release();
  return CATT_OK;
}

void catt_ctx_destroy(catt_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  c->V.release(); c->J.release();
  free_align_buffers(c->ab[0]); free_align_buffers(c->ab[1]);
  cudaFree(c->d_finder); cudaFree(c->d_dseq); cudaFree(c->d_doff);
  for (auto& ev : c->ev) if (ev) cudaEventDestroy(ev);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

int catt_ref_count(const catt_ctx* c, int which) { return c ? (int)(which ? c->J : c->V).names.size() : 0; }
const char* catt_ref_name(const catt_ctx* c, int which, int32_t id) {
  if (!c) return nullptr;
  const RefSet& rs = which ? c->J : c->V;
  return (id >= 0 && id < (int)rs.names.size()) ? rs.names[id].c_str() : nullptr;
}
int catt_ref_seq(const catt_ctx* c, int which, int32_t id, char* out, int32_t cap) {
  if (!c) return CATT_E_ARG;
  const RefSet& rs = which ? c->J : c->V;
  if (id < 0 || id >= (int)rs.names.size()) { set_error("record id out of range"); return CATT_E_ARG; }
  const int l = rs.len[id];
  for (int i = 0; i < l && i < cap; i++) out[i] = "ACGT"[rs.T[rs.off[id] + i]];
  return l;
}
uint64_t catt_ctx_stream(const catt_ctx* c) { return c ? (uint64_t)(uintptr_t)c->stream : 0; }

int catt_run_reads(catt_ctx* c, int64_t n, const char* names, const int64_t* name_off, const char* seqs, const int64_t* seq_off,
                   const char* quals, catt_result** out) {
  if (!c || !out || n < 0 || (n > 0 && (!names || !name_off || !seqs || !seq_off))) { set_error("catt_run_reads: bad argument"); return CATT_E_ARG; }
  CATT_CUDA(cudaSetDevice(c->device));
  HostReads hr;
  hr.n = n; hr.names = names; hr.name_off = name_off; hr.seqs = seqs; hr.seq_off = seq_off; hr.quals = quals;
  return run_all(c, hr, out);
}

int catt_run_file(catt_ctx* c, const char* path, catt_result** out) {
  if (!c || !path || !out) { set_error("catt_run_file: null argument"); return CATT_E_ARG; }
  CATT_CUDA(cudaSetDevice(c->device));
  HostReads hr;
  int rc = parse_fastx(path, &hr);
  if (rc) return rc;
  return run_all(c, hr, out);
}

int catt_result_info(const catt_result* r, catt_info* out) { if (!r || !out) return CATT_E_ARG; *out = r->info; return CATT_OK; }
int catt_result_reads(const catt_result* r, int which, const catt_read** reads, int64_t* n) {
  if (!r || !reads || !n || which < 0 || which > 1) return CATT_E_ARG;
  *reads = r->part[which].data(); *n = (int64_t)r->part[which].size();
  return CATT_OK;
}
int catt_align_records(const catt_result* r, int which, const catt_aln** recs, int64_t* n) {
  if (!r || !recs || !n || which < 0 || which > 1) return CATT_E_ARG;
  *recs = r->rec[which].data(); *n = (int64_t)r->rec[which].size();
  return CATT_OK;
}
void catt_result_free(catt_result* r) { delete r; }

int catt_best_d(catt_ctx* c, const char* const* cdr3, int64_t n, int32_t* out) {
  if (!c || !out || (n > 0 && !cdr3)) return CATT_E_ARG;
  if (c->d_seqs.empty()) { for (int64_t i = 0; i < n; i++) out[i] = -1; return CATT_OK; }
  CATT_CUDA(cudaSetDevice(c->device));
  std::string cat; std::vector<int32_t> off{0};
  for (int64_t i = 0; i < n; i++) { cat += cdr3[i]; off.push_back((int32_t)cat.size()); }
  char* d_c = nullptr; int32_t* d_o = nullptr; int32_t* d_r = nullptr;
  CATT_CUDA(cudaMalloc(&d_c, cat.size() + 16));
  CATT_CUDA(cudaMalloc(&d_o, off.size() * sizeof(int32_t)));
  CATT_CUDA(cudaMalloc(&d_r, std::max<int64_t>(n, 1) * sizeof(int32_t)));
  CATT_CUDA(cudaMemcpyAsync(d_c, cat.data(), cat.size(), cudaMemcpyHostToDevice, c->stream));
  CATT_CUDA(cudaMemcpyAsync(d_o, off.data(), off.size() * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
  int rc = launch_best_d(d_c, d_o, n, c->d_dseq, c->d_doff, (int)c->d_seqs.size(), d_r, c->stream);
  if (!rc) {
    cudaError_t e = cudaMemcpyAsync(out, d_r, n * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) rc = cuda_fail(e, "best_d", __FILE__, __LINE__);
  }
  cudaFree(d_c); cudaFree(d_o); cudaFree(d_r);
  return rc;
}

// ---- stage-level ---------------------------------------------------------------------------------------------------
int catt_batch_upload(catt_ctx* c, int64_t n, const char* seqs, const int64_t* seq_off, int64_t first_ordinal, catt_batch** out) {
  if (!c || !out || n < 0 || (n > 0 && (!seqs || !seq_off))) { set_error("catt_batch_upload: bad argument"); return CATT_E_ARG; }
  CATT_CUDA(cudaSetDevice(c->device));
  PackedReads pr;
  int rc = pack_reads(n, seqs, seq_off, &pr);
  if (rc) return rc;
  std::unique_ptr<catt_batch> b(new catt_batch());
  b->first_ordinal = first_ordinal;
  if ((rc = upload_reads(pr, n, &b->dr, c->stream))) return rc;
  CATT_CUDA(cudaStreamSynchronize(c->stream));
  *out = b.release();
  return CATT_OK;
}

int catt_batch_align(catt_ctx* c, catt_batch* b) {
  if (!c || !b) return CATT_E_ARG;
  CATT_CUDA(cudaSetDevice(c->device));
  b->info = catt_info{};
  b->info.n_reads = b->dr.n;
  int launches = 0, rc;
  // the J records must survive the V pass and vice versa: each reference set has its own buffers
  for (int which = 0; which < 2; which++)
    if ((rc = align_refset(c, which, b->dr.view(), b->first_ordinal, &b->info, &launches))) return rc;
  b->info.kernel_launches = launches;
  return CATT_OK;
}

int catt_batch_sync(catt_ctx* c, catt_batch* b, catt_info* info) {
  if (!c || !b) return CATT_E_ARG;
  CATT_CUDA(cudaStreamSynchronize(c->stream));
  if (info) *info = b->info;
  return CATT_OK;
}

int catt_batch_download(catt_ctx* c, catt_batch* b, catt_aln* v_out, catt_aln* j_out) {
  if (!c || !b) return CATT_E_ARG;
  if (v_out) CATT_CUDA(cudaMemcpy(v_out, c->ab[0].recs, b->dr.n * sizeof(catt_aln), cudaMemcpyDeviceToHost));
  if (j_out) CATT_CUDA(cudaMemcpy(j_out, c->ab[1].recs, b->dr.n * sizeof(catt_aln), cudaMemcpyDeviceToHost));
  return CATT_OK;
}

void catt_batch_free(catt_batch* b) { if (b) { b->dr.release(); delete b; } }

int catt_align_shard(catt_ctx* c, int64_t n, const char* seqs, const int64_t* seq_off, int64_t first_ordinal, catt_aln* v_out,
                     catt_aln* j_out) {
  catt_batch* b = nullptr;
  int rc = catt_batch_upload(c, n, seqs, seq_off, first_ordinal, &b);
  if (rc) return rc;
  rc = catt_batch_align(c, b);
  if (!rc) rc = catt_batch_download(c, b, v_out, j_out);
  catt_batch_free(b);
  return rc;
}

int catt_assemble(catt_ctx* c, int64_t n, const char* names, const int64_t* name_off, const char* seqs, const int64_t* seq_off,
                  const char* quals, const catt_aln* v_recs, const catt_aln* j_recs, catt_result** out) {
  if (!c || !out || !v_recs || !j_recs || n < 0) { set_error("catt_assemble: bad argument"); return CATT_E_ARG; }
  CATT_CUDA(cudaSetDevice(c->device));
  HostReads hr;
  hr.n = n; hr.names = names; hr.name_off = name_off; hr.seqs = seqs; hr.seq_off = seq_off; hr.quals = quals;
  std::unique_ptr<catt_result> R(new catt_result());
  R->info.n_reads = n;
  R->rec[0].assign(v_recs, v_recs + n); R->rec[1].assign(j_recs, j_recs + n);
  for (int64_t i = 0; i < n; i++) { R->info.v_hits += v_recs[i].mapped; R->info.j_hits += j_recs[i].mapped; }
  PackedReads pr;
  int rc = pack_reads(n, seqs, seq_off, &pr);
  if (rc) return rc;
  DeviceReads dr;
  if ((rc = upload_reads(pr, n, &dr, c->stream))) return rc;
  CATT_CUDA(cudaStreamSynchronize(c->stream));
  int launches = 0;
  rc = stage_b(c, hr, dr, R.get(), &launches);
  dr.release();
  if (rc) return rc;
  R->info.kernel_launches = launches;
  *out = R.release();
  return CATT_OK;
}

// ---- parity hooks ----------------------------------------------------------------------------------------------------
int catt_seed_candidates(catt_ctx* c, int which, int64_t n, const char* seqs, const int64_t* seq_off, uint8_t* cand_out) {
  if (!c || !cand_out || which < 0 || which > 1) return CATT_E_ARG;
  CATT_CUDA(cudaSetDevice(c->device));
  PackedReads pr;
  int rc = pack_reads(n, seqs, seq_off, &pr);
  if (rc) return rc;
  DeviceReads dr;
  if ((rc = upload_reads(pr, n, &dr, c->stream))) return rc;
  RefSet& rs = which ? c->J : c->V;
  AlignBuffers& ab = c->ab[which];
  int launches = 0;
  if ((rc = ensure_align_buffers(ab, rs.dev, n))) return rc;
  CATT_CUDA(cudaMemsetAsync(ab.counters, 0, 8 * sizeof(unsigned long long), c->stream));
  if ((rc = launch_seed(rs.dev, dr.view(), ab, c->sm_count, c->stream, &launches))) return rc;
  std::vector<uint32_t> bits((size_t)n * rs.dev.CW);
  CATT_CUDA(cudaMemcpyAsync(bits.data(), ab.candbits, bits.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  CATT_CUDA(cudaStreamSynchronize(c->stream));
  dr.release();
  const int nb = rs.dev.nrec * 2;
  for (int64_t i = 0; i < n; i++)
    for (int b = 0; b < nb; b++) cand_out[i * nb + b] = (bits[i * rs.dev.CW + (b >> 5)] >> (b & 31)) & 1u;
  return CATT_OK;
}

int catt_sw_pairs(catt_ctx* c, int which, int64_t n, const char* seqs, const int64_t* seq_off, const int32_t* rid, const int32_t* strand,
                  int32_t* out3) {
  if (!c || !out3 || !rid || !strand || which < 0 || which > 1) return CATT_E_ARG;
  CATT_CUDA(cudaSetDevice(c->device));
  RefSet& rs = which ? c->J : c->V;
  for (int64_t i = 0; i < n; i++) if (rid[i] < 0 || rid[i] >= (int)rs.names.size()) { set_error("rid out of range"); return CATT_E_ARG; }
  PackedReads pr;
  int rc = pack_reads(n, seqs, seq_off, &pr);
  if (rc) return rc;
  DeviceReads dr;
  if ((rc = upload_reads(pr, n, &dr, c->stream))) return rc;
  int32_t *d_rid = nullptr, *d_strand = nullptr, *d_out = nullptr;
  CATT_CUDA(cudaMalloc(&d_rid, std::max<int64_t>(n, 1) * 4)); CATT_CUDA(cudaMalloc(&d_strand, std::max<int64_t>(n, 1) * 4));
  CATT_CUDA(cudaMalloc(&d_out, std::max<int64_t>(n, 1) * 12));
  CATT_CUDA(cudaMemcpyAsync(d_rid, rid, n * 4, cudaMemcpyHostToDevice, c->stream));
  CATT_CUDA(cudaMemcpyAsync(d_strand, strand, n * 4, cudaMemcpyHostToDevice, c->stream));
  rc = launch_sw_pairs(rs.dev, dr.view(), d_rid, d_strand, d_out, c->sm_count, c->stream);
  if (!rc) {
    cudaError_t e = cudaMemcpy(out3, d_out, n * 12, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = cuda_fail(e, "sw_pairs D2H", __FILE__, __LINE__);
  }
  cudaFree(d_rid); cudaFree(d_strand); cudaFree(d_out); dr.release();
  return rc;
}

int catt_finder(catt_ctx* c, int array_version, int64_t n, const char* seqs, const int64_t* seq_off, int32_t* out3) {
  if (!c || !out3) return CATT_E_ARG;
  CATT_CUDA(cudaSetDevice(c->device));
  PackedReads pr;
  int rc = pack_reads(n, seqs, seq_off, &pr);
  if (rc) return rc;
  DeviceReads dr;
  if ((rc = upload_reads(pr, n, &dr, c->stream))) return rc;
  int32_t* d_out = nullptr;
  CATT_CUDA(cudaMalloc(&d_out, std::max<int64_t>(n, 1) * 12));
  rc = launch_finder_raw(dr.view(), c->d_finder, array_version, d_out, c->stream);
  if (!rc) {
    cudaError_t e = cudaStreamSynchronize(c->stream);
    if (e == cudaSuccess) e = cudaMemcpy(out3, d_out, n * 12, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = cuda_fail(e, "finder D2H", __FILE__, __LINE__);
  }
  cudaFree(d_out); dr.release();
  return rc;
}

int catt_alu_peak(catt_ctx* c, double* ops, double* ms) {
  if (!c || !ops || !ms) return CATT_E_ARG;
  CATT_CUDA(cudaSetDevice(c->device));
  return alu_peak(c->sm_count, c->stream, ops, ms);
}

}  // extern "C"

// api.cu — the C ABI of libcatt_b200 (include/catt_b200.h): context, Stage A driver (map2align on the GPU), FASTQ input,
// stage-level and parity entry points.  Stage B lives in assemble.cu.
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

#include <fcntl.h>
#include <omp.h>
#include <unistd.h>

#include "gz.h"
#include "host_types.h"

namespace catt {
const char* last_error();
int parse_bam(const char* path, HostReads* hr);          // bam.cu
}  // namespace catt

using namespace catt;

namespace {

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
  if (!ab.counters) CATT_CUDA(cudaMalloc(&ab.counters, N_COUNTERS * sizeof(unsigned long long)));
  if (n <= ab.cap_reads) return CATT_OK;
  cudaFree(ab.candbits); cudaFree(ab.canonbits); cudaFree(ab.ntask); cudaFree(ab.task_off); cudaFree(ab.gapped_list);
  cudaFree(ab.wtasks); cudaFree(ab.wres); cudaFree(ab.worder); cudaFree(ab.whist);
  ab.cap_reads = n + n / 8 + 64;
  CATT_CUDA(cudaMalloc(&ab.candbits, ab.cap_reads * ref.CW * sizeof(uint32_t)));
  CATT_CUDA(cudaMalloc(&ab.canonbits, ab.cap_reads * ref.CW * sizeof(uint32_t)));
  CATT_CUDA(cudaMalloc(&ab.ntask, (ab.cap_reads + 1) * sizeof(uint32_t)));
  CATT_CUDA(cudaMalloc(&ab.task_off, (ab.cap_reads + 1) * sizeof(uint32_t)));
  CATT_CUDA(cudaMalloc(&ab.gapped_list, ab.cap_reads * sizeof(uint32_t)));
  CATT_CUDA(cudaMalloc(&ab.wtasks, ab.cap_reads * sizeof(SwTask)));
  CATT_CUDA(cudaMalloc(&ab.wres, ab.cap_reads * 2 * sizeof(uint32_t)));
  CATT_CUDA(cudaMalloc(&ab.worder, (ab.cap_reads + MAX_REC + 2) * sizeof(uint32_t)));
  CATT_CUDA(cudaMalloc(&ab.whist, (2 * MAX_REC + 4) * sizeof(uint32_t)));
  return CATT_OK;
}

void free_align_buffers(AlignBuffers& ab) {
  cudaFree(ab.candbits); cudaFree(ab.canonbits); cudaFree(ab.ntask); cudaFree(ab.task_off); cudaFree(ab.tasks); cudaFree(ab.task_score);
  cudaFree(ab.gapped_list); cudaFree(ab.wtasks); cudaFree(ab.wres); cudaFree(ab.worder); cudaFree(ab.whist); cudaFree(ab.counters); cudaFree(ab.cub_tmp);
  cudaFree(ab.seed_scratch);
  ab = AlignBuffers{};
}

// Stage A (map2align) of one reference set for one chunk of reads, queued on the stream without a host round trip: task
// counts, winner counts and the gapped list are consumed on the device.  align_finish() reads the counters after the
// caller's stream sync; if the task buffer turned out too small the chunk is re-run once with the exact size.
// The J reference set runs on its own stream: its kernels are small and fill the tails of the V kernels (and vice versa); both wait
// for the same inputs (ev_inputs, recorded on the main stream once the chunk's reads are packed).
inline cudaStream_t aln_stream(catt_ctx* c, int which) { return which ? c->stream_j : c->stream; }
int fork_j(catt_ctx* c) {
  CATT_CUDA(cudaEventRecord(c->ev_inputs, c->stream));
  CATT_CUDA(cudaStreamWaitEvent(c->stream_j, c->ev_inputs, 0));
  return CATT_OK;
}
int join_streams(catt_ctx* c) {
  CATT_CUDA(cudaStreamSynchronize(c->stream));
  CATT_CUDA(cudaStreamSynchronize(c->stream_j));
  return CATT_OK;
}

int align_tail(catt_ctx* c, int which, const ReadsDev& reads, catt_aln* recs, int64_t first_ordinal, int* launches) {
  RefSet& rs = which ? c->J : c->V;
  AlignBuffers& ab = c->ab[which];
  cudaStream_t st = aln_stream(c, which);
  cudaEvent_t* ev = c->ev_a[which];
  int rc;
  CATT_CUDA(cudaEventRecord(ev[2], st));
  if ((rc = launch_expand(rs.dev, reads, ab, st, launches))) return rc;
  CATT_CUDA(cudaEventRecord(ev[3], st));
  if ((rc = launch_sw(rs.dev, reads, ab, c->sm_count, st, launches))) return rc;
  CATT_CUDA(cudaEventRecord(ev[4], st));
  if ((rc = launch_select(rs.dev, reads, ab, recs, first_ordinal, c->min_score, c->sm_count, st, launches))) return rc;
  CATT_CUDA(cudaEventRecord(ev[5], st));
  CATT_CUDA(cudaMemcpyAsync(c->pin[PIN_CNT].as<unsigned long long>() + which * N_COUNTERS, ab.counters, N_COUNTERS * sizeof(unsigned long long),
                            cudaMemcpyDeviceToHost, st));
  return CATT_OK;
}

int align_queue(catt_ctx* c, int which, const ReadsDev& reads, catt_aln* recs, int64_t first_ordinal, int* launches) {
  RefSet& rs = which ? c->J : c->V;
  AlignBuffers& ab = c->ab[which];
  cudaStream_t st = aln_stream(c, which);
  int rc;
  if (reads.n == 0) return CATT_OK;
  if ((rc = ensure_align_buffers(ab, rs.dev, reads.n))) return rc;
  CATT_CUDA(cudaMemsetAsync(ab.counters, 0, N_COUNTERS * sizeof(unsigned long long), st));
  CATT_CUDA(cudaMemsetAsync(ab.ntask, 0, (reads.n + 1) * sizeof(uint32_t), st));
  CATT_CUDA(cudaEventRecord(c->ev_a[which][0], st));
  if ((rc = launch_seed(rs.dev, reads, ab, c->sm_count, st, launches))) return rc;
  CATT_CUDA(cudaEventRecord(c->ev_a[which][1], st));
  return align_tail(c, which, reads, recs, first_ordinal, launches);
}

int align_finish(catt_ctx* c, int which, const ReadsDev& reads, catt_aln* recs, int64_t first_ordinal, catt_info* info, int* launches) {
  if (reads.n == 0) return CATT_OK;
  AlignBuffers& ab = c->ab[which];
  cudaStream_t st = aln_stream(c, which);
  cudaEvent_t* ev = c->ev_a[which];
  unsigned long long* cnt = c->pin[PIN_CNT].as<unsigned long long>() + which * N_COUNTERS;
  float ms_seed = 0, ms_exp = 0, ms_sw = 0, ms_sel = 0;
  cudaEventElapsedTime(&ms_seed, ev[0], ev[1]);
  int rc;
  for (int attempt = 0;; attempt++) {
    if (cnt[6] > 0) { set_error("%llu reads exceed the seeding run-list capacity", cnt[6]); return CATT_E_LIMIT; }
    float a = 0, b = 0, d = 0;
    cudaEventElapsedTime(&a, ev[2], ev[3]);
    cudaEventElapsedTime(&b, ev[3], ev[4]);
    cudaEventElapsedTime(&d, ev[4], ev[5]);
    ms_exp += a; ms_sw += b; ms_sel += d;
    if (cnt[9] == 0) break;
    if (attempt == 1) { set_error("SW task buffer overflow persists (%llu tasks)", cnt[9]); return CATT_E_LIMIT; }
    if ((rc = ensure_task_buffers(ab, (int64_t)cnt[9] + (int64_t)cnt[9] / 8 + 1024))) return rc;
    unsigned long long z[N_COUNTERS] = {0};
    z[4] = cnt[4]; z[6] = cnt[6];
    CATT_CUDA(cudaMemcpyAsync(ab.counters, z, sizeof z, cudaMemcpyHostToDevice, st));
    CATT_CUDA(cudaStreamSynchronize(st));
    if ((rc = align_tail(c, which, reads, recs, first_ordinal, launches))) return rc;
    CATT_CUDA(cudaStreamSynchronize(st));
  }
  if (info) {
    // the stage times are those of the V reference set (its stream carries > 95 % of the work); the J set runs concurrently on its
    // own stream and is reported as one interval
    if (which == 0) { info->ms_seed += ms_seed + ms_exp; info->ms_seed_kernel += ms_seed; info->ms_sw += ms_sw; info->ms_select += ms_sel; }
    else info->ms_j_stream += ms_seed + ms_exp + ms_sw + ms_sel;
    if (which) { info->cand_j += (int64_t)cnt[0]; info->cells_j += (int64_t)cnt[1]; info->gapped_j += (int64_t)cnt[2]; info->j_hits += (int64_t)cnt[3]; info->cells_computed_j += (int64_t)cnt[8]; }
    else { info->cand_v += (int64_t)cnt[0]; info->cells_v += (int64_t)cnt[1]; info->gapped_v += (int64_t)cnt[2]; info->v_hits += (int64_t)cnt[3]; info->cells_computed_v += (int64_t)cnt[8]; }
  }
  return CATT_OK;
}

// Size the per-chunk Stage A scratch of both reference sets for the largest chunk that will come, before the pipeline is busy:
// growing it chunk by chunk along the size ramp costs a cudaFree (device-wide sync) + cudaMalloc per step on the first call.
int presize_align(catt_ctx* c, int64_t total_reads) {
  const int64_t maxchunk = std::min<int64_t>(total_reads, c->chunk_reads * 3 / 2);
  if (maxchunk <= 0) return CATT_OK;
  int rc;
  for (int which = 0; which < 2; which++) {
    RefSet& rs = which ? c->J : c->V;
    if ((rc = ensure_align_buffers(c->ab[which], rs.dev, maxchunk)) || (rc = ensure_task_buffers(c->ab[which], maxchunk * 10 + 4096))) return rc;
  }
  return CATT_OK;
}

int ensure_recs(catt_ctx* c, int64_t n) {
  int rc;
  for (int which = 0; which < 2; which++)
    if ((rc = c->recs_all[which].ensure((size_t)std::max<int64_t>(n, 1) * sizeof(catt_aln)))) return rc;
  return CATT_OK;
}

// Stage A over reads that are already resident: chunk by chunk (bounded scratch), records into ctx->recs_all.  `fh->start`
// lists the input files: chunks never straddle a file and the read ordinal restarts at every file.
int stage_a_range(catt_ctx* c, const DeviceReads& dr, int64_t lo, int64_t hi, int64_t ordinal0, int file, FileHits* fh, catt_info* info,
                  int* launches) {
  int rc;
  for (int64_t i0 = lo; i0 < hi;) {
    const int64_t left = hi - i0;
    const int64_t cnt = left <= c->chunk_reads * 3 / 2 ? left : c->chunk_reads;      // no tiny tail chunk
    const int64_t ord = ordinal0 + (i0 - lo);
    const ReadsDev rv = dr.view(i0, cnt);
    if ((rc = fork_j(c))) return rc;
    for (int which = 0; which < 2; which++)
      if ((rc = align_queue(c, which, rv, c->recs_all[which].as<catt_aln>() + i0, ord, launches))) return rc;
    if ((rc = join_streams(c))) return rc;
    const int64_t v0 = info->v_hits, j0 = info->j_hits;
    for (int which = 0; which < 2; which++)
      if ((rc = align_finish(c, which, rv, c->recs_all[which].as<catt_aln>() + i0, ord, info, launches))) return rc;
    fh->v[file] += info->v_hits - v0; fh->j[file] += info->j_hits - j0;
    i0 += cnt;
  }
  return CATT_OK;
}

int stage_a_resident(catt_ctx* c, const DeviceReads& dr, int64_t first_ordinal, FileHits* fh, catt_info* info, int* launches,
                     const std::vector<int64_t>* file_ordinal0 = nullptr) {
  int rc;
  if ((rc = ensure_recs(c, dr.n)) || (rc = presize_align(c, dr.n))) return rc;
  const int nfiles = (int)fh->start.size() - 1;
  fh->v.assign(nfiles, 0); fh->j.assign(nfiles, 0);
  for (int f = 0; f < nfiles; f++)
    if ((rc = stage_a_range(c, dr, fh->start[f], fh->start[f + 1], file_ordinal0 ? (*file_ordinal0)[f] : first_ordinal, f, fh, info, launches))) return rc;
  return CATT_OK;
}

// Host reads -> ctx->rd_* on the device, Stage A pipelined: while the kernels of chunk i run, the host threads pack chunk
// i+1 into pinned staging and the copy stream moves it to HBM.  `fh->start` lists the input files of the sample: chunks never
// straddle a file and bwa's read ordinal restarts at every file (each file is one `bwa mem` run, catt.jl:143-147).
int stage_a_stream(catt_ctx* c, const HostReads& hr, int64_t first_ordinal, bool align, bool want_quals, DeviceReads* view, DeviceQuals* dq,
                   FileHits* fh, catt_info* info, int* launches, const std::function<int()>& under_kernels = nullptr,
                   const std::vector<int64_t>* file_ordinal0 = nullptr) {
  const int64_t n = hr.n;
  int rc;
  const double t0 = now_ms();
  c->shard.valid = false;                          // the read buffers are about to be overwritten (catt_shard_upload re-validates)
  if ((rc = c->pin[PIN_OFF].ensure((size_t)(n + 1) * sizeof(uint32_t))) || (rc = c->pin[PIN_LEN].ensure((size_t)std::max<int64_t>(n, 1) * sizeof(int32_t)))) return rc;
  uint32_t* h_off = c->pin[PIN_OFF].as<uint32_t>();
  int32_t* h_len = c->pin[PIN_LEN].as<int32_t>();
  int max_len = 0;
  if ((rc = pack_layout(n, hr.seq_off, h_off, h_len, &max_len))) return rc;
  const uint64_t total_words = h_off[n];
  if ((rc = c->rd_words.ensure((total_words + 4) * sizeof(uint32_t))) || (rc = c->rd_off.ensure((size_t)(n + 1) * sizeof(uint32_t))) ||
      (rc = c->rd_len.ensure((size_t)std::max<int64_t>(n, 1) * sizeof(int32_t))) || (rc = ensure_recs(c, n)) || (align && (rc = presize_align(c, n)))) return rc;
  DeviceReads dr;
  dr.words = c->rd_words.as<uint32_t>(); dr.off = c->rd_off.as<uint32_t>(); dr.len = c->rd_len.as<int32_t>(); dr.n = n; dr.max_len = max_len;
  *view = dr;
  cudaStream_t st = c->stream, cs = c->copy_stream;
  // the quality strings follow their chunk to HBM on the copy stream (the result strings are written on the device)
  char* d_quals = nullptr;
  if (dq) *dq = DeviceQuals{};
  if ((rc = c->rd_seqoff.ensure((size_t)(n + 1) * sizeof(int64_t)))) return rc;
  const int64_t* d_seqoff = c->rd_seqoff.as<int64_t>();
  // the per-read offset / length tables travel with their chunk (stage_chunk) instead of in one unhidden copy up front
  if (want_quals && dq && n > 0 && c->only_cdr3 && !c->resident_quals) {
    dq->seq_off = d_seqoff;                       // the strings follow later, for the reads that keep a CDR3 only (assemble.cu)
    dq->host_quals = hr.quals; dq->host_seq_off = hr.seq_off;
  } else if (want_quals && dq && n > 0) {
    dq->seq_off = d_seqoff;
    if (hr.quals) {
      // seq_off may be a view into a larger sample (one rank's block of a GPU group): the device copy holds this block only and is
      // addressed with the caller's absolute offsets
      if ((rc = c->rd_quals.ensure((size_t)(hr.seq_off[n] - hr.seq_off[0]) + 16))) return rc;
      d_quals = c->rd_quals.as<char>() - hr.seq_off[0];
      dq->quals = d_quals;
    }
  }
  if (info) info->bytes_h2d += (int64_t)((n + 1) * 4 + n * 4 + (n + 1) * 8);
  struct Chunk { int64_t i0, cnt, ordinal; int file; };
  std::vector<Chunk> chunks;
  const int nfiles = (int)fh->start.size() - 1;
  fh->v.assign(nfiles, 0); fh->j.assign(nfiles, 0);
  // the first chunk is small (nothing hides its copy, the GPU should get work early) and the sizes ramp up x4 per chunk: the
  // copy of chunk i+1 always fits under the kernels of chunk i
  int64_t ramp = std::max<int64_t>(c->chunk_reads / 64, 1024);
  for (int f = 0; f < nfiles; f++)
    for (int64_t i0 = fh->start[f]; i0 < fh->start[f + 1];) {
      const int64_t want = std::min<int64_t>(ramp, c->chunk_reads);
      ramp = std::min<int64_t>(ramp * 4, c->chunk_reads);
      const int64_t cnt = std::min<int64_t>(want, fh->start[f + 1] - i0);
      chunks.push_back(Chunk{i0, cnt, (file_ordinal0 ? (*file_ordinal0)[f] : first_ordinal) + (i0 - fh->start[f]), f});
      i0 += cnt;
    }
  const int64_t nchunks = (int64_t)chunks.size();
  // the host only moves bytes: the raw bases (and quality strings) of chunk ci are copied by all host threads into pinned staging
  // and follow on the copy stream; the 2-bit packing runs on the device in front of the chunk's kernels
  auto par_copy = [](char* dst, const char* src, size_t bytes) {
    const int64_t piece = 1 << 20, np = (int64_t)((bytes + piece - 1) / piece);
    #pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < np; p++) memcpy(dst + p * piece, src + p * piece, (size_t)std::min<int64_t>(piece, (int64_t)bytes - p * piece));
  };
  auto stage_chunk = [&](int64_t ci) -> int {
    const int64_t i0 = chunks[ci].i0, i1 = i0 + chunks[ci].cnt;
    const size_t bytes = (size_t)(hr.seq_off[i1] - hr.seq_off[i0]);
    if (ci >= 2) CATT_CUDA(cudaEventSynchronize(c->ev_q[ci & 1]));        // the copies that last read these staging buffers
    PinBuf& pb = c->pin[PIN_PACK0 + (ci & 1)];
    int r;
    if ((r = pb.ensure(bytes + 16)) || (r = c->rd_raw[ci & 1].ensure(bytes + 16))) return r;
    par_copy(pb.as<char>(), hr.seqs + hr.seq_off[i0], bytes);
    CATT_CUDA(cudaMemcpyAsync(dr.off + i0, h_off + i0, (size_t)(i1 - i0 + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, cs));
    CATT_CUDA(cudaMemcpyAsync(dr.len + i0, h_len + i0, (size_t)(i1 - i0) * sizeof(int32_t), cudaMemcpyHostToDevice, cs));
    CATT_CUDA(cudaMemcpyAsync(c->rd_seqoff.as<int64_t>() + i0, hr.seq_off + i0, (size_t)(i1 - i0 + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, cs));
    CATT_CUDA(cudaMemcpyAsync(c->rd_raw[ci & 1].p, pb.p, bytes, cudaMemcpyHostToDevice, cs));
    CATT_CUDA(cudaEventRecord(c->ev_h2d[ci & 1], cs));
    if (info) info->bytes_h2d += (int64_t)bytes;
    if (d_quals && bytes) {
      PinBuf& qb = c->pin[PIN_QUAL0 + (ci & 1)];
      if ((r = qb.ensure(bytes + 16))) return r;
      par_copy(qb.as<char>(), hr.quals + hr.seq_off[i0], bytes);
      CATT_CUDA(cudaMemcpyAsync(d_quals + hr.seq_off[i0], qb.p, bytes, cudaMemcpyHostToDevice, cs));
      if (info) info->bytes_h2d += (int64_t)bytes;
    }
    CATT_CUDA(cudaEventRecord(c->ev_q[ci & 1], cs));
    return CATT_OK;
  };
  if (nchunks > 0 && (rc = stage_chunk(0))) return rc;
  if (info) info->ms_pack += now_ms() - t0;
  catt_info local{};
  int64_t queued_reads = 0;
  bool hook_done = false;
  const bool trace_chunks = getenv("CATT_TRACE_CHUNKS") != nullptr;
  double tt_launch = 0, tt_stage = 0, tt_join = 0, tt_finish = 0;
  for (int64_t ci = 0; ci < nchunks; ci++) {
    const double tc0 = now_ms();
    const Chunk ch = chunks[ci];
    const ReadsDev rv = dr.view(ch.i0, ch.cnt);
    CATT_CUDA(cudaStreamWaitEvent(st, c->ev_h2d[ci & 1], 0));
    if ((rc = launch_pack_raw(c->rd_raw[ci & 1].as<char>() - hr.seq_off[ch.i0], d_seqoff + ch.i0, dr.len + ch.i0, dr.off + ch.i0, ch.cnt, dr.words,
                              c->sm_count, st, launches))) return rc;
    if (align) {
      if ((rc = fork_j(c))) return rc;
      for (int which = 0; which < 2; which++)
        if ((rc = align_queue(c, which, rv, c->recs_all[which].as<catt_aln>() + ch.i0, ch.ordinal, launches))) return rc;
    }
    const double tc1 = now_ms();
    if (ci + 1 < nchunks && (rc = stage_chunk(ci + 1))) return rc;
    const double tc2 = now_ms();
    queued_reads += ch.cnt;
    if (under_kernels && !hook_done && (queued_reads >= 200000 || ci + 1 == nchunks)) {      // host work hidden under this chunk's kernels
      hook_done = true;
      if ((rc = under_kernels())) return rc;
    }
    if ((rc = join_streams(c))) return rc;
    const double tc3 = now_ms();
    if (align) {
      const int64_t v0 = local.v_hits, j0 = local.j_hits;
      for (int which = 0; which < 2; which++)
        if ((rc = align_finish(c, which, rv, c->recs_all[which].as<catt_aln>() + ch.i0, ch.ordinal, &local, launches))) return rc;
      fh->v[ch.file] += local.v_hits - v0; fh->j[ch.file] += local.j_hits - j0;
    }
    tt_launch += tc1 - tc0; tt_stage += tc2 - tc1; tt_join += tc3 - tc2; tt_finish += now_ms() - tc3;
  }
  if (trace_chunks)
    fprintf(stderr, "[catt trace] stage_a_stream: %lld chunks; host ms: launch %.1f, stage next %.1f, wait for the GPU %.1f, counters %.1f\n",
            (long long)nchunks, tt_launch, tt_stage, tt_join, tt_finish);
  CATT_CUDA(cudaStreamSynchronize(st));
  CATT_CUDA(cudaStreamSynchronize(cs));
  c->last_n = n; c->last_words = (int64_t)total_words; c->last_max_len = max_len;
  if (under_kernels && !hook_done && (rc = under_kernels())) return rc;
  if (info && align) {
    info->ms_seed += local.ms_seed; info->ms_seed_kernel += local.ms_seed_kernel; info->ms_sw += local.ms_sw; info->ms_select += local.ms_select;
    info->ms_j_stream += local.ms_j_stream;
    info->cand_v += local.cand_v; info->cand_j += local.cand_j; info->cells_v += local.cells_v; info->cells_j += local.cells_j;
    info->gapped_v += local.gapped_v; info->gapped_j += local.gapped_j; info->v_hits += local.v_hits; info->j_hits += local.j_hits;
    info->cells_computed_v += local.cells_computed_v; info->cells_computed_j += local.cells_computed_j;
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
  if (hr->own_name_off.empty()) { hr->own_name_off.push_back(0); hr->own_seq_off.push_back(0); }   // appends when called per input file
  size_t b, e;
  while (line(b, e)) {
    size_t nb = b;
    while (nb < e && (txt[nb] == ' ' || txt[nb] == '\t')) nb++;
    if (nb == e) continue;                               // blank line (kseq skips to the next '@' / '>')
    const char tag = txt[b];
    if (tag != '@' && tag != '>') { set_error("%s: malformed FASTA/FASTQ record near byte %zu", path.c_str(), b); return CATT_E_IO; }
    size_t k = b + 1;
    while (k < e && txt[k] != ' ' && txt[k] != '\t') k++;
    hr->own_names.append(txt, b + 1, k - b - 1);
    hr->own_name_off.push_back((int64_t)hr->own_names.size());
    size_t sb, se;
    if (!line(sb, se)) { set_error("%s: truncated record", path.c_str()); return CATT_E_IO; }
    hr->own_seqs.append(txt, sb, se - sb);
    if (tag == '>') {                                    // multi-line FASTA: the sequence runs to the next header (kseq)
      while (p < N && txt[p] != '>') {
        size_t cb, ce;
        if (!line(cb, ce)) break;
        hr->own_seqs.append(txt, cb, ce - cb);
        se += ce - cb;
      }
    }
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

int download_records(catt_ctx* c, int64_t n, catt_result* R) {
  for (int which = 0; which < 2; which++) {
    R->rec[which].resize(n);
    const double td = now_ms();
    if (n) CATT_CUDA(cudaMemcpy(R->rec[which].data(), c->recs_all[which].p, n * sizeof(catt_aln), cudaMemcpyDeviceToHost));
    R->info.ms_d2h += now_ms() - td;
  }
  return CATT_OK;
}

// ---- catt_run_file(s) without host parsing: the raw FASTQ text goes to HBM and ingest.cu indexes, validates and packs it there.
// Returns CATT_OK with *handled = false when the input is not strict 4-line FASTQ (FASTA, wrapped or blank lines, ...): the host
// parser then takes over (and owns the error messages).
bool read_whole(const char* path, std::string* gz_out, int64_t* plain_size, bool* is_gz) {
  FILE* f = fopen(path, "rb");
  if (!f) return false;
  unsigned char magic[2] = {0, 0};
  const size_t got = fread(magic, 1, 2, f);
  *is_gz = got == 2 && magic[0] == 0x1f && magic[1] == 0x8b;
  if (!*is_gz) { fseeko(f, 0, SEEK_END); *plain_size = (int64_t)ftello(f); }
  fclose(f);
  if (*is_gz) return slurp_file(path, *gz_out);
  return true;
}

int run_files_device(catt_ctx* c, int nfiles, const char* const* paths, catt_result** out, bool* handled, bool bam = false) {
  *handled = false;
  Trace tr;
  std::vector<std::string> gz((size_t)nfiles);
  std::vector<int64_t> size((size_t)nfiles, 0), tpos((size_t)nfiles + 1, 0);
  std::vector<char> isgz((size_t)nfiles, 0);
  // single-member .gz files are inflated wave by wave straight into the pinned text (pgz.cu), so the first pieces cross PCIe and
  // are aligned while the rest of the file is still being inflated; their size is the gzip trailer's ISIZE (verified at the end)
  std::vector<RawFile> raw((size_t)nfiles);
  std::vector<std::unique_ptr<PgzStream>> pgz((size_t)nfiles);
  std::vector<std::unique_ptr<BgzfStream>> bgzf((size_t)nfiles);
  std::vector<std::unique_ptr<BamTextStream>> bamtx((size_t)nfiles);
  std::vector<std::function<int(char*, size_t, size_t*)>> src((size_t)nfiles);      // streamed sources: next(dst, cap, &wrote)
  for (int f = 0; f < nfiles; f++) {
    bool g = false;
    {
      FILE* fp = fopen(paths[f], "rb");
      unsigned char magic[2] = {0, 0};
      if (!fp) { set_error("cannot read %s", paths[f]); return CATT_E_IO; }
      g = fread(magic, 1, 2, fp) == 2 && magic[0] == 0x1f && magic[1] == 0x8b;
      fclose(fp);
    }
    if (bam && !g) return CATT_OK;                                   // SAM text / uncompressed BAM: the host decoder (bam.cu)
    if (g && !getenv("CATT_ZLIB_ONLY")) {
      if (!raw[f].open(paths[f])) { set_error("cannot read %s", paths[f]); return CATT_E_IO; }
      const size_t rn = raw[f].size();
      int64_t bound = -1;
      if (bam) {                                                     // BGZF BAM -> FASTQ text, group by group (bam.cu)
        std::unique_ptr<BamTextStream> bt(new BamTextStream());
        if (!bt->open(raw[f].data(), rn)) return CATT_OK;            // gzipped SAM, damaged chain: the host decoder reports it
        bound = (int64_t)bt->text_bound();
        bamtx[f] = std::move(bt);
        BamTextStream* q = bamtx[f].get();
        src[f] = [q](char* d, size_t cap, size_t* w) { return q->next(d, cap, w); };
      } else {
        const char* cb = getenv("CATT_PGZ_CHUNK");
        std::unique_ptr<PgzStream> st(new PgzStream());
        std::unique_ptr<BgzfStream> bs(new BgzfStream());
        const bool opened = (cb || rn >= (1u << 20)) && st->open(raw[f].data(), rn, cb ? (size_t)atol(cb) : (1u << 20));
        // the trailer's ISIZE is the text length mod 2^32: plausible (at least half the compressed size) -> exact bound; a large file
        // with an implausible one has wrapped -> generous bound (8x; FASTQ compresses 3-6x), HBM only - the host side is a window
        // ... and so has a large file whose ISIZE passes that test but gives a ratio no FASTQ reaches (text of 4 GiB and more: 9.6 GB of
        // reads leave ISIZE ~ 1 GB beside a 1.9 GB .gz).  Plain FASTQ deflates 3-6x; below 2.5x on a file of 512 MB and more the field
        // is taken as wrapped
        const uint64_t isz = opened ? (uint64_t)st->isize_field() : 0;
        const bool wrapped = opened && rn >= (1u << 29) && (isz * 2 < rn || isz * 2 < rn * 5);
        if (opened && ((isz > 0 && isz * 2 >= rn) || wrapped)) {
          bound = wrapped ? (int64_t)rn * 8 : (int64_t)st->isize_field();   // the stream verifies ISIZE (mod 2^32) and the CRC-32 at the end
          pgz[f] = std::move(st);
          PgzStream* q = pgz[f].get();
          src[f] = [q](char* d, size_t cap, size_t* w) { return q->next(d, cap, w); };
        } else if (bs->open(raw[f].data(), rn) && bs->total() > 0) { // bgzip'd FASTQ: independent members, exact size known
          bound = (int64_t)bs->total();
          bgzf[f] = std::move(bs);
          BgzfStream* q = bgzf[f].get();
          src[f] = [q](char* d, size_t cap, size_t* w) { return q->next(d, cap, w); };
        }
      }
      if (bound >= 0) {
        size[f] = bound;
        isgz[f] = 2;
        tpos[f + 1] = (tpos[f] + size[f] + 16 + 15) & ~(int64_t)15;
        continue;
      }
      raw[f].close();
    }
    if (!read_whole(paths[f], &gz[f], &size[f], &g)) { set_error("cannot read %s", paths[f]); return CATT_E_IO; }
    isgz[f] = g;
    if (g) size[f] = (int64_t)gz[f].size();
    tpos[f + 1] = tpos[f] + size[f] + 16;                          // room for the terminating newline, 16-byte aligned starts
    tpos[f + 1] = (tpos[f + 1] + 15) & ~(int64_t)15;
  }
  int rc;
  const int64_t piece_max_full = std::max<int64_t>(c->chunk_reads * 320, 1 << 20);
  // The text stays in HBM for Stage B, but on the host a piece is dead as soon as its copy has completed: inputs larger than the
  // window pass through a sliding pinned window instead of being pinned whole.  Pinned memory costs ~1 ms per MB (allocation + first
  // touch) and a process that handles one sample - the Julia driver - pays it inside its only call, so the window is small: 320 MB
  // with pieces of at most 80 MB (Stage A is as efficient on 265 k-read pieces as on 1 M-read ones; measured steady state 149.6 ms
  // per 2 M reads against 144.6 ms with 768 MB, first call 375 against 856 ms).  The parallel gzip decoder produces up to a third of
  // the window per wave and gets 640 MB.  CATT_PIN_WINDOW overrides (tests).
  bool any_pgz = false;
  for (int f = 0; f < nfiles; f++) any_pgz = any_pgz || (bool)pgz[f];
  int64_t wcap = any_pgz ? (640ll << 20) : (320ll << 20);
  if (const char* e = getenv("CATT_PIN_WINDOW")) wcap = std::max<int64_t>(atoll(e), 1 << 20);
  const int64_t piece_max = std::min<int64_t>(piece_max_full, std::max<int64_t>(wcap / 4, 1 << 20));
  const bool windowed = tpos[nfiles] + 16 > wcap;
  if (windowed) wcap = std::max<int64_t>(wcap, 3 * piece_max);            // tail of an unfinished piece + the next block / wave
  if (c->pin[PIN_TEXT].ensure((size_t)(windowed ? wcap : tpos[nfiles]) + 16, true) != CATT_OK) { cudaGetLastError(); return CATT_OK; }   // cannot pin that much: host path
  if ((rc = c->rd_text.ensure((size_t)tpos[nfiles] + 16))) return rc;
  char* h = c->pin[PIN_TEXT].as<char>();
  char* d_text = c->rd_text.as<char>();
  // A reader thread brings the files into pinned memory block by block and, as soon as a record-aligned piece is complete (a line
  // that starts with '@' whose second next line starts with '+'), queues its copy to HBM on the copy stream and publishes it:
  // piece k+1 is read and crosses PCIe while piece k is indexed, packed and aligned.  Pieces ramp up x4 from a small first one.
  struct Piece { int64_t t0, t1; int file; };
  struct Feed {
    std::mutex mu; std::condition_variable cv;
    std::vector<Piece> pieces; bool done = false; int status = CATT_OK; bool fallback = false; std::string err;
  } feed;
  {
    int64_t total = 0;
    for (int f = 0; f < nfiles; f++) total += size[f];
    const size_t want = (size_t)(16 + nfiles * 8 + total / piece_max * 2);
    const size_t old = c->ev_piece.size();
    if (old < want) {
      c->ev_piece.resize(want);
      for (size_t k = old; k < want; k++) CATT_CUDA(cudaEventCreateWithFlags(&c->ev_piece[k], cudaEventDisableTiming));
    }
  }
  const int device = c->device;
  std::thread reader([&, device]() {
    cudaSetDevice(device);
    auto fail = [&](int code, const std::string& msg, bool fb) {
      std::lock_guard<std::mutex> lk(feed.mu);
      feed.status = code; feed.err = msg; feed.fallback = fb; feed.done = true;
      feed.cv.notify_all();
    };
    int64_t piece_bytes = std::max<int64_t>(piece_max / 64, 1 << 20);
    size_t n_published = 0;
    for (int f = 0; f < nfiles; f++) {
      // dst + x is where byte x of file f lives on the host.  Whole-input mode: h + tpos[f] + x.  Windowed: h + (x - wbase), where
      // wbase = the first byte still held (everything before the next piece start p has been published; the window slides once
      // those copies have completed)
      // (a view, not a pointer: in windowed mode byte 0 of the file may lie before the buffer)
      struct FileView {
        char* base; int64_t origin;                              // base[0] holds byte `origin` of the file
        char* operator+(int64_t x) const { return base + (x - origin); }
        char& operator[](int64_t x) const { return base[x - origin]; }
      };
      auto offset_of = [](const char* p, const FileView& v) -> int64_t { return (int64_t)(p - v.base) + v.origin; };
      int64_t wbase = 0;
      FileView dst{windowed ? h : h + tpos[f], 0};
      int64_t have = 0, end = -1, p = 0;                           // bytes of the file produced so far; trimmed end once known; next piece start
      if (windowed && n_published && cudaEventSynchronize(c->ev_piece[n_published - 1]) != cudaSuccess) { fail(CATT_E_CUDA, "FASTQ text upload failed", false); return; }
      auto make_room = [&](int64_t need) -> bool {                 // room for bytes [have, have + need) of this file
        if (!windowed || (have - wbase) + need + 16 <= wcap) return true;
        if (n_published && cudaEventSynchronize(c->ev_piece[n_published - 1]) != cudaSuccess) return false;
        memmove(h, dst + p, (size_t)(have - p));
        if (getenv("CATT_TRACE_WINDOW")) fprintf(stderr, "[catt trace] pinned window slides: file %d, origin %lld -> %lld, %lld bytes kept\n", f, (long long)wbase, (long long)p, (long long)(have - p));
        wbase = p;
        dst.origin = wbase;
        return (have - wbase) + need + 16 <= wcap;
      };
      int fd = -1;
      int64_t gz_copied = 0;                                       // isgz == 1: bytes of the inflated string copied into the window
      if (isgz[f] == 0 && (fd = open(paths[f], O_RDONLY)) < 0) { fail(CATT_E_IO, std::string("cannot read ") + paths[f], false); return; }
      while (true) {
        if (have < size[f] && isgz[f] == 1) {                      // inflated beforehand (zlib path): copy it in, block by block
          const int64_t blk = std::min<int64_t>(size[f] - have, windowed ? std::min<int64_t>(256 << 20, wcap / 4) : size[f]);
          if (!make_room(blk)) { fail(CATT_OK, "", true); return; }
          memcpy(dst + have, gz[f].data() + gz_copied, (size_t)blk);
          gz_copied += blk; have += blk;
          if (have == size[f]) std::string().swap(gz[f]);
        } else if (have < size[f] && isgz[f] == 2) {               // next wave of the parallel inflate
          size_t wrote = 0;
          if (!make_room(std::min<int64_t>(size[f] - have, wcap / 3))) { fail(CATT_OK, "", true); return; }
          const int64_t cap = windowed ? std::min<int64_t>(size[f] - have, wcap - 16 - (have - wbase)) : size[f] - have;
          const int r = src[f](dst + have, (size_t)cap, &wrote);
          have += (int64_t)wrote;
          // a failed check anywhere in the stream (or a size bound that does not hold): nothing it produced may be trusted - the
          // host path (zlib / the host BAM decoder) takes the whole sample
          if (r < 0) { fail(CATT_OK, "", true); return; }
          if (r == 0) size[f] = have;                                // the source is exhausted: now the size is exact
          else if (wrote == 0) continue;
        } else if (have < size[f]) {                               // next block, all host threads
          const int64_t blk = std::min<int64_t>(size[f] - have, std::min<int64_t>(32 << 20, std::max<int64_t>(wcap / 8, 1 << 16))), sub = 4 << 20, ns = (blk + sub - 1) / sub;
          if (!make_room(blk)) { if (fd >= 0) close(fd); fail(CATT_OK, "", true); return; }
          bool ok = true;
          #pragma omp parallel for schedule(dynamic, 1) reduction(&& : ok)
          for (int64_t k = 0; k < ns; k++) {
            int64_t o = have + k * sub;
            const int64_t e = std::min(have + blk, o + sub);
            while (o < e) {
              const ssize_t r = pread(fd, dst + o, (size_t)(e - o), (off_t)o);
              if (r <= 0) { ok = false; break; }
              o += r;
            }
          }
          if (!ok) { close(fd); fail(CATT_E_IO, std::string("short read on ") + paths[f], false); return; }
          have += blk;
        }
        if (have == size[f] && end < 0) {                          // whole file present: trim trailing blank space, terminate with '\n'
          int64_t e = size[f];
          while (e > wbase && (dst[e - 1] == '\n' || dst[e - 1] == '\r' || dst[e - 1] == ' ' || dst[e - 1] == '\t')) e--;
          if (e == 0) { if (fd >= 0) close(fd); fail(CATT_OK, "", true); return; }
          dst[e++] = '\n';
          end = e;
        }
        if (p == 0 && have > 0 && dst[0] != '@') { if (fd >= 0) close(fd); fail(CATT_OK, "", true); return; }   // not FASTQ: host path
        // cut as many pieces as the bytes at hand allow
        while (true) {
          const int64_t limit = end >= 0 ? end : have;
          int64_t q = p + piece_bytes;
          if (end >= 0 && q >= end) q = end;
          else {
            if (q + 65536 > limit && end < 0) break;               // the boundary search needs a little look-ahead
            bool found = false;
            while (q < limit) {                                    // advance q to the start of a header line
              const char* nl = (const char*)memchr(dst + q, '\n', (size_t)(limit - q));
              if (!nl) break;
              q = offset_of(nl, dst) + 1;
              if (q >= limit) break;
              if (dst[q] != '@') continue;
              const char* n1 = (const char*)memchr(dst + q, '\n', (size_t)(limit - q));
              const char* n2 = n1 ? (const char*)memchr(n1 + 1, '\n', (size_t)((dst + limit) - (n1 + 1))) : nullptr;
              if (n2 && n2 + 1 < dst + limit && n2[1] == '+') { found = true; break; }
            }
            if (!found) { if (end >= 0) q = end; else break; }
          }
          if (q <= p) break;
          const Piece pc{tpos[f] + p, tpos[f] + q, f};
          if (cudaMemcpyAsync(d_text + pc.t0, dst + p, (size_t)(pc.t1 - pc.t0), cudaMemcpyHostToDevice, c->copy_stream) != cudaSuccess ||
              n_published >= c->ev_piece.size() || cudaEventRecord(c->ev_piece[n_published], c->copy_stream) != cudaSuccess) {
            if (fd >= 0) close(fd);
            fail(CATT_E_CUDA, "FASTQ text upload failed", false);
            return;
          }
          {
            std::lock_guard<std::mutex> lk(feed.mu);
            feed.pieces.push_back(pc);
          }
          feed.cv.notify_all();
          n_published++;
          p = q;
          piece_bytes = std::min(piece_bytes * 4, piece_max);
          if (end >= 0 && p >= end) break;
        }
        if (end >= 0 && p >= end) break;
      }
      if (fd >= 0) close(fd);
    }
    std::lock_guard<std::mutex> lk(feed.mu);
    feed.done = true;
    feed.cv.notify_all();
  });
  struct Joiner { std::thread& t; ~Joiner() { if (t.joinable()) t.join(); } } joiner{reader};
  auto next_piece = [&](size_t k, Piece* out) -> int {          // 1 = piece k available, 0 = no more pieces, < 0 = error / fallback
    std::unique_lock<std::mutex> lk(feed.mu);
    feed.cv.wait(lk, [&] { return feed.pieces.size() > k || feed.done; });
    if (feed.pieces.size() > k) { *out = feed.pieces[k]; return 1; }
    if (feed.fallback) return -2;
    if (feed.status != CATT_OK) { set_error("%s", feed.err.c_str()); return -1; }
    return 0;
  };
  tr.lap("run_files: read text + queue H2D");
  std::unique_ptr<catt_result> R(new catt_result());
  int launches = 0;
  FileHits fh;
  fh.start = {0};
  fh.v.assign(nfiles, 0); fh.j.assign(nfiles, 0);
  int64_t n = 0;
  uint64_t words = 0;
  int max_len = 0;
  // the seeding kernel sizes its scratch by the longest read of the sample, which is only known at the end: use the limit of the
  // first piece's reads for that piece, then the running maximum (a later, longer read only enlarges later launches)
  int prev_file = -1;
  // Indexing + packing of piece k+1 (three small host round trips on the ingest stream) runs while Stage A of piece k occupies the
  // alignment streams: queue Stage A of k, ingest k+1, then wait for Stage A of k.
  struct Ingested { Piece pc{0, 0, 0}; int64_t np = 0; uint64_t wp = 0; double ms = 0; };
  int64_t n_ing = 0; uint64_t words_ing = 0;              // reads / words ingested so far (n, words: through Stage A)
  bool want_fallback = false; int ing_rc = CATT_OK;
  auto fetch_and_ingest = [&](size_t k, Ingested* g) -> int {       // 1 = piece k ingested, 0 = no more pieces, -1 = want_fallback / ing_rc
    const int got = next_piece(k, &g->pc);
    if (got == 0) return 0;
    if (got == -2) { want_fallback = true; return -1; }            // not strict FASTQ: the host parser takes over
    if (got < 0) { ing_rc = feed.status; return -1; }
    const double t0 = now_ms();
    if (cudaStreamWaitEvent(c->stream_ingest, c->ev_piece[k], 0) != cudaSuccess) { ing_rc = CATT_E_CUDA; set_error("ingest stream wait failed"); return -1; }
    int fb = 0;
    g->np = 0; g->wp = 0;
    const int r = ingest_fastq(c, d_text, g->pc.t0, g->pc.t1, n_ing, words_ing, &g->np, &g->wp, &max_len, &fb, &launches, c->stream_ingest);
    if (r) { if (fb) want_fallback = true; else ing_rc = r; return -1; }
    if (cudaEventRecord(c->ev_ingest, c->stream_ingest) != cudaSuccess) { ing_rc = CATT_E_CUDA; set_error("ingest event failed"); return -1; }
    n_ing += g->np; words_ing += g->wp;
    g->ms = now_ms() - t0;
    return 1;
  };
  Ingested cur, nxt;
  int more = fetch_and_ingest(0, &cur);
  if (more < 0) return want_fallback ? CATT_OK : ing_rc;
  for (size_t k = 0; more == 1; k++) {
    const Piece& pc = cur.pc;
    const int64_t np = cur.np;
    if (prev_file >= 0 && prev_file != pc.file) fh.start.push_back(n);
    prev_file = pc.file;
    R->info.bytes_h2d += pc.t1 - pc.t0;
    // nothing is in flight on the alignment streams here; the ingest stream may still be packing piece k
    if (k == 0 && np > 0 && tpos[nfiles] > 4 * (pc.t1 - pc.t0)) {
      // The first piece tells bytes and packed words per read: size the per-read buffers for the whole input once instead of growing
      // them piece by piece (each growth is a cudaMalloc + copy + cudaFree: ~2 s of the first call on a 9.5 GB file)
      CATT_CUDA(cudaStreamSynchronize(c->stream_ingest));
      const double per_read = (double)(pc.t1 - pc.t0) / (double)np;
      const int64_t est = (int64_t)((double)tpos[nfiles] / per_read * 1.03) + 4096;
      const uint64_t est_words = (uint64_t)((double)cur.wp / (double)np * (double)est) + 4096;
      cudaStream_t st = c->stream;
      if (est_words < 0xFFFFFFF0ull &&
          ((rc = c->rd_name_beg.ensure_keep((size_t)(est + 1) * 8, st)) || (rc = c->rd_name_end.ensure_keep((size_t)(est + 1) * 8, st)) ||
           (rc = c->rd_seq_beg.ensure_keep((size_t)(est + 1) * 8, st)) || (rc = c->rd_qual_beg.ensure_keep((size_t)(est + 1) * 8, st)) ||
           (rc = c->rd_len.ensure_keep((size_t)(est + 1) * 4, st)) || (rc = c->rd_off.ensure_keep((size_t)(est + 2) * 4, st)) ||
           (rc = c->rd_words.ensure_keep((size_t)(est_words + 4) * 4, st)) ||
           (rc = c->recs_all[0].ensure_keep((size_t)est * sizeof(catt_aln), st)) || (rc = c->recs_all[1].ensure_keep((size_t)est * sizeof(catt_aln), st))))
        return rc;
      if ((rc = presize_align(c, est))) return rc;
    }
    for (int which = 0; which < 2; which++)
      if ((rc = c->recs_all[which].ensure_keep((size_t)std::max<int64_t>(n + np, 1) * sizeof(catt_aln), c->stream))) return rc;
    CATT_CUDA(cudaStreamWaitEvent(c->stream, c->ev_ingest, 0));   // the packed reads of piece k are complete
    const int64_t file_first = fh.start[pc.file];
    const int64_t ord = n - file_first;
    const int piece_max_len = max_len;                             // as of piece k (the next ingest may raise it)
    auto reads_now = [&]() {                                       // current addresses of the read buffers (a growth moves them)
      DeviceReads dr;
      dr.words = c->rd_words.as<uint32_t>(); dr.off = c->rd_off.as<uint32_t>(); dr.len = c->rd_len.as<int32_t>(); dr.n = n + np; dr.max_len = piece_max_len;
      return dr;
    };
    const double ta = now_ms();
    if (np > 0 && np <= c->chunk_reads * 3 / 2) {                  // one Stage A chunk: queue it, ingest the next piece under it, then finish
      const ReadsDev rv = reads_now().view(n, np);
      if ((rc = fork_j(c))) return rc;
      for (int which = 0; which < 2; which++)
        if ((rc = align_queue(c, which, rv, c->recs_all[which].as<catt_aln>() + n, ord, &launches))) { join_streams(c); return rc; }
      more = fetch_and_ingest(k + 1, &nxt);
      if ((rc = join_streams(c))) return rc;
      if (more < 0) return want_fallback ? CATT_OK : ing_rc;
      const ReadsDev rv2 = reads_now().view(n, np);
      const int64_t v0 = R->info.v_hits, j0 = R->info.j_hits;
      for (int which = 0; which < 2; which++)
        if ((rc = align_finish(c, which, rv2, c->recs_all[which].as<catt_aln>() + n, ord, &R->info, &launches))) return rc;
      fh.v[pc.file] += R->info.v_hits - v0; fh.j[pc.file] += R->info.j_hits - j0;
    } else {                                                       // several chunks (short reads): the plain loop, then the next piece
      if (np > 0 && (rc = stage_a_range(c, reads_now(), n, n + np, ord, pc.file, &fh, &R->info, &launches))) return rc;
      more = fetch_and_ingest(k + 1, &nxt);
      if (more < 0) return want_fallback ? CATT_OK : ing_rc;
    }
    if (tr.on) fprintf(stderr, "[catt trace]   piece %zu: %lld bytes, %lld reads, ingest %.1f ms (under the previous piece's Stage A), stage A %.1f ms\n", k,
                       (long long)(pc.t1 - pc.t0), (long long)np, cur.ms, now_ms() - ta);
    n += np; words += cur.wp;
    cur = nxt;
  }
  fh.start.push_back(n);
  if ((int)fh.start.size() != nfiles + 1) { set_error("run_files: piece bookkeeping"); return CATT_E_ARG; }
  tr.lap("run_files: index + pack + align, piece by piece");
  R->info.n_reads = n;
  DeviceReads dr;
  dr.words = c->rd_words.as<uint32_t>(); dr.off = c->rd_off.as<uint32_t>(); dr.len = c->rd_len.as<int32_t>(); dr.n = n; dr.max_len = max_len;
  if (c->keep_records && (rc = download_records(c, n, R.get()))) return rc;
  NamesDev nd;
  nd.s = d_text; nd.beg = reinterpret_cast<const int64_t*>(c->rd_name_beg.p); nd.end = reinterpret_cast<const int64_t*>(c->rd_name_end.p);
  DeviceQuals dq;
  dq.quals = d_text; dq.seq_off = reinterpret_cast<const int64_t*>(c->rd_qual_beg.p);
  if ((rc = stage_b(c, n, nd, dr.view(), dq, fh, R.get(), &launches))) return rc;
  tr.lap("run_files: stage_b");
  R->info.kernel_launches = launches;
  *out = R.release();
  *handled = true;
  return CATT_OK;
}

int run_all(catt_ctx* c, const HostReads& hr, const std::vector<int64_t>& file_start, catt_result** out) {
  std::unique_ptr<catt_result> R(new catt_result());
  int launches = 0;
  Trace tr;
  R->info.n_reads = hr.n;
  DeviceReads dr;
  DeviceQuals dq;
  FileHits fh;
  fh.start = file_start;
  NamesDev nd;
  catt_result* Rp = R.get();
  // the QNAMEs (pageable, ~20 B per read) cross PCIe while the kernels of a large chunk run
  int rc = stage_a_stream(c, hr, 0, true, true, &dr, &dq, &fh, &R->info, &launches,
                          [&]() { return upload_names(c, hr.n, hr.names, hr.name_off, &nd, &Rp->info, c->copy_stream); });
  tr.lap("run_all: pack + H2D + align (pipelined)");
  if (rc == CATT_OK && c->keep_records) rc = download_records(c, hr.n, R.get());
  if (rc == CATT_OK) rc = stage_b(c, hr.n, nd, dr.view(), dq, fh, R.get(), &launches);
  tr.lap("run_all: stage_b");
  if (rc) return rc;
  R->info.kernel_launches = launches;
  *out = R.release();
  return CATT_OK;
}

// ---- one sample over several GPUs -------------------------------------------------------------------------------------
// Contiguous block [lo, hi) of rank r: sizes differ by at most one read, earlier ranks take the extra ones.
void block_bounds(int64_t n, int nranks, int r, int64_t* lo, int64_t* hi) {
  const int64_t base = n / nranks, extra = n % nranks;
  *lo = r * base + std::min<int64_t>(r, extra);
  *hi = *lo + base + (r < extra ? 1 : 0);
}

// This block's share of every input file; bwa's read ordinal restarts with every file (catt.jl:143-147).
int rank_files(int64_t n_total, int64_t lo, int64_t n_loc, const std::vector<int64_t>& gstart, FileHits* fh, std::vector<int64_t>* ord0) {
  const int nfiles = (int)gstart.size() - 1;
  if (nfiles < 1 || gstart.front() != 0 || gstart.back() != n_total || lo < 0 || lo + n_loc > n_total) {
    set_error("GPU group: block [%lld, %lld) outside the sample of %lld reads", (long long)lo, (long long)(lo + n_loc), (long long)n_total);
    return CATT_E_ARG;
  }
  fh->start = {0};
  ord0->clear();
  for (int f = 0; f < nfiles; f++) {
    const int64_t a = std::min(std::max(gstart[f], lo), lo + n_loc), b = std::min(std::max(gstart[f + 1], lo), lo + n_loc);
    fh->start.push_back(b - lo);
    ord0->push_back(a - gstart[f]);
  }
  return CATT_OK;
}

int rank_finish(catt_ctx* c, Comm* comm, int root, int64_t n_total, int64_t lo, int64_t n_loc, const std::vector<int64_t>& gstart, const NamesDev& nd,
                const DeviceQuals& dq, int max_len, double ms_a, int rc, std::unique_ptr<catt_result>& R, int launches, catt_result** out);

// One rank: Stage A on this rank's block `hr` (reads [lo, lo + hr.n) of the sample), then its part of the global Stage B.
// `gstart` are the global file boundaries.  Every rank returns a result; the lists are on `root`.
int run_rank(catt_ctx* c, Comm* comm, int root, int64_t n_total, int64_t lo, const HostReads& hr, const std::vector<int64_t>& gstart, catt_result** out) {
  std::unique_ptr<catt_result> R(new catt_result());
  int launches = 0, rc;
  Trace tr;
  FileHits fh;
  std::vector<int64_t> ord0;
  if ((rc = rank_files(n_total, lo, hr.n, gstart, &fh, &ord0))) return rc;
  DeviceReads dr;
  DeviceQuals dq;
  NamesDev nd;
  catt_result* Rp = R.get();
  const double t0 = now_ms();
  rc = stage_a_stream(c, hr, 0, true, true, &dr, &dq, &fh, &R->info, &launches,
                      [&]() { return upload_names(c, hr.n, hr.names, hr.name_off, &nd, &Rp->info, c->copy_stream); }, &ord0);
  const double ms_a = now_ms() - t0;
  tr.lap("run_rank: pack + H2D + align (pipelined)");
  return rank_finish(c, comm, root, n_total, lo, hr.n, gstart, nd, dq, dr.max_len, ms_a, rc, R, launches, out);
}

// Everything after this rank's Stage A: the first exchange (block bounds + status), the global Stage B, the merged counters.
int rank_finish(catt_ctx* c, Comm* comm, int root, int64_t n_total, int64_t lo, int64_t n_loc, const std::vector<int64_t>& gstart, const NamesDev& nd,
                const DeviceQuals& dq, int max_len, double ms_a, int rc, std::unique_ptr<catt_result>& R, int launches, catt_result** out) {
  Trace tr;
  const int N = comm->nranks;
  // a rank that failed must still answer the first exchange, or its peers would wait for it forever: the status travels with the bounds
  long long mine[4] = {lo, n_loc, rc ? -1 : (long long)max_len, rc};
  std::vector<long long> all((size_t)4 * N);
  int rc2 = comm->host_allgather(mine, all.data(), sizeof mine, c->stream);
  if (rc) return rc;
  if (rc2) return rc2;
  ShardIn in;
  in.comm = comm; in.lo = lo; in.n_loc = n_loc; in.n_words = c->last_words; in.gstart = gstart; in.root = root;
  in.bounds.assign((size_t)N + 1, 0);
  int64_t expect = 0;
  for (int r = 0; r < N; r++) {
    if (all[4 * r + 3] != 0) { set_error("GPU group: rank %d failed in Stage A (code %lld)", r, all[4 * r + 3]); return (int)all[4 * r + 3]; }
    if (all[4 * r] != expect) { set_error("GPU group: the blocks of the ranks are not contiguous (rank %d starts at %lld, expected %lld)", r, all[4 * r], (long long)expect); return CATT_E_ARG; }
    in.bounds[r] = expect; expect += all[4 * r + 1];
    in.max_len = std::max(in.max_len, (int)all[4 * r + 2]);
  }
  in.bounds[N] = expect;
  if (expect != n_total) { set_error("GPU group: the blocks cover %lld reads, the sample has %lld", (long long)expect, (long long)n_total); return CATT_E_ARG; }
  if (c->keep_records) {                            // diagnostics: every rank's records, in input order, on the root
    std::vector<size_t> cn(N), dp(N);
    for (int r = 0; r < N; r++) { cn[r] = (size_t)(in.bounds[r + 1] - in.bounds[r]) * sizeof(catt_aln); dp[r] = (size_t)in.bounds[r] * sizeof(catt_aln); }
    for (int which = 0; which < 2; which++) {
      if ((rc = c->sh[SH_SLOTS - 1].ensure((size_t)std::max<int64_t>(n_total, 1) * sizeof(catt_aln)))) return rc;
      if ((rc = comm->allgatherv(c->recs_all[which].p, c->sh[SH_SLOTS - 1].p, cn.data(), dp.data(), c->stream))) return rc;
      CATT_CUDA(cudaStreamSynchronize(c->stream));
      if (comm->rank == root) {
        R->rec[which].resize((size_t)n_total);
        if (n_total) CATT_CUDA(cudaMemcpy(R->rec[which].data(), c->sh[SH_SLOTS - 1].p, (size_t)n_total * sizeof(catt_aln), cudaMemcpyDeviceToHost));
      }
    }
  }
  if ((rc = stage_b_sharded(c, in, nd, dq, R.get(), &launches))) return rc;
  tr.lap("rank_finish: stage_b_sharded");
  // counters: sums over the ranks; stage times: the slowest rank
  catt_info& I = R->info;
  I.kernel_launches = launches;
  I.ms_stage_a_max = I.ms_stage_a_min = ms_a;
  std::vector<catt_info> infos((size_t)N);
  if ((rc = comm->host_allgather(&I, infos.data(), sizeof(catt_info), c->stream))) return rc;
  catt_info T = infos[(size_t)root];                 // the globally derived values (lists, k, the_ERR, ...) are identical on every rank
  auto sum = [&](int64_t catt_info::*m) { int64_t v = 0; for (auto& x : infos) v += x.*m; T.*m = v; };
  auto mx = [&](double catt_info::*m) { double v = 0; for (auto& x : infos) v = std::max(v, x.*m); T.*m = v; };
  sum(&catt_info::v_hits); sum(&catt_info::j_hits); sum(&catt_info::cand_v); sum(&catt_info::cand_j); sum(&catt_info::cells_v); sum(&catt_info::cells_j);
  sum(&catt_info::gapped_v); sum(&catt_info::gapped_j); sum(&catt_info::cells_computed_v); sum(&catt_info::cells_computed_j);
  sum(&catt_info::kernel_launches); sum(&catt_info::bytes_h2d); sum(&catt_info::bytes_d2h); sum(&catt_info::rescued); sum(&catt_info::bytes_comm);
  mx(&catt_info::ms_seed); mx(&catt_info::ms_sw); mx(&catt_info::ms_select); mx(&catt_info::ms_extract); mx(&catt_info::ms_pool); mx(&catt_info::ms_h2d);
  mx(&catt_info::ms_d2h); mx(&catt_info::ms_host); mx(&catt_info::ms_order); mx(&catt_info::ms_pack); mx(&catt_info::ms_seed_kernel);
  mx(&catt_info::ms_pool_kernel); mx(&catt_info::ms_strings); mx(&catt_info::ms_j_stream); mx(&catt_info::ms_comm); mx(&catt_info::ms_stage_a_max);
  for (int k = 0; k < 8; k++) {
    double m = 0; int64_t b = 0;
    for (auto& x : infos) { m = std::max(m, x.ms_coll[k]); b += x.bytes_coll[k]; }
    T.ms_coll[k] = m; T.bytes_coll[k] = b;
  }
  T.ms_stage_a_min = infos[0].ms_stage_a_min;
  for (auto& x : infos) T.ms_stage_a_min = std::min(T.ms_stage_a_min, x.ms_stage_a_min);
  T.n_reads = n_total; T.n_ranks = N;
  I = T;
  *out = R.release();
  return CATT_OK;
}

// ---- a FASTQ file on a GPU group: every rank reads ITS byte range of the file (cut at record starts), moves the text to its GPU,
// indexes / validates / packs it there (ingest.cu) and learns how many reads it holds; one exchange of those counts gives every
// rank the global ordinal of its first read (bwa's tie-break hashes it), then Stage A runs on the resident block and Stage B is the
// global one.  N PCIe links and N shares of the host cores work in parallel; nothing is parsed on the host.
// Returns RANK_FALLBACK (on every rank) when the file is not strict 4-line FASTQ: the caller takes the host parser.
constexpr int RANK_FALLBACK = 1;

// first record start at or after byte x: a line that starts with '@' whose second next line starts with '+'
int64_t record_boundary(int fd, int64_t fsize, int64_t x) {
  if (x <= 0) return 0;
  if (x >= fsize) return fsize;
  std::vector<char> buf;
  for (size_t span = 1 << 18; span <= ((size_t)1 << 28); span <<= 2) {
    const int64_t from = x - 1;                                  // one byte back: is x itself a line start?
    const size_t want = (size_t)std::min<int64_t>((int64_t)span, fsize - from);
    buf.resize(want);
    size_t got = 0;
    while (got < want) { const ssize_t r = pread(fd, buf.data() + got, want - got, (off_t)(from + (int64_t)got)); if (r <= 0) break; got += (size_t)r; }
    if (got < want) return -1;
    const char* b = buf.data();
    const char* end = b + got;
    const char* p = b;
    while (p < end) {
      const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
      if (!nl || nl + 1 >= end) break;
      const char* q = nl + 1;                                    // a line start
      if (*q == '@') {
        const char* n1 = (const char*)memchr(q, '\n', (size_t)(end - q));
        const char* n2 = n1 ? (const char*)memchr(n1 + 1, '\n', (size_t)(end - (n1 + 1))) : nullptr;
        if (n2 && n2 + 1 < end) { if (n2[1] == '+') return from + (q - b); }
        else if (from + (int64_t)got < fsize) break;             // need more look-ahead
      }
      p = q;
    }
    if (from + (int64_t)got >= fsize) return fsize;              // no further record start: the tail belongs to the previous block
  }
  return -1;
}

// ---- Stage A in two halves, for inputs whose global read ordinals are not known yet (run_file_rank) ------------------------------
// front: seeding, task expansion, SW scores, selection up to the tie-break (best score, candidate count, bitmap of the candidates
// that reach it - per read, kept for the whole block in c->df_*); back: the tie-break with the ordinal, TRACK pass, CIGAR facts.
int deferred_reserve(catt_ctx* c, int64_t n, cudaStream_t st) {
  int rc;
  for (int which = 0; which < 2; which++) {
    const int CW = (which ? c->J : c->V).dev.CW;
    if ((rc = c->df_tied[which].ensure_keep((size_t)std::max<int64_t>(n, 1) * CW * sizeof(uint32_t), st)) ||
        (rc = c->df_mx[which].ensure_keep((size_t)std::max<int64_t>(n, 1) * sizeof(int16_t), st)) ||
        (rc = c->df_nc[which].ensure_keep((size_t)std::max<int64_t>(n, 1) * sizeof(int16_t), st))) return rc;
  }
  return CATT_OK;
}

int front_launch(catt_ctx* c, int which, const ReadsDev& rv, int64_t r0, bool with_seed, int* launches) {
  RefSet& rs = which ? c->J : c->V;
  AlignBuffers& ab = c->ab[which];
  cudaStream_t st = aln_stream(c, which);
  cudaEvent_t* ev = c->ev_a[which];
  int rc;
  if (with_seed) {
    CATT_CUDA(cudaMemsetAsync(ab.counters, 0, N_COUNTERS * sizeof(unsigned long long), st));
    CATT_CUDA(cudaMemsetAsync(ab.ntask, 0, (rv.n + 1) * sizeof(uint32_t), st));
    CATT_CUDA(cudaEventRecord(ev[0], st));
    if ((rc = launch_seed(rs.dev, rv, ab, c->sm_count, st, launches))) return rc;
    CATT_CUDA(cudaEventRecord(ev[1], st));
  }
  CATT_CUDA(cudaEventRecord(ev[2], st));
  if ((rc = launch_expand(rs.dev, rv, ab, st, launches))) return rc;
  CATT_CUDA(cudaEventRecord(ev[3], st));
  if ((rc = launch_sw(rs.dev, rv, ab, c->sm_count, st, launches))) return rc;
  CATT_CUDA(cudaEventRecord(ev[4], st));
  if ((rc = launch_select_pre(rs.dev, rv, ab, c->df_tied[which].as<uint32_t>() + r0 * rs.dev.CW, c->df_mx[which].as<int16_t>() + r0,
                              c->df_nc[which].as<int16_t>() + r0, st, launches))) return rc;
  CATT_CUDA(cudaEventRecord(ev[5], st));
  CATT_CUDA(cudaMemcpyAsync(c->pin[PIN_CNT].as<unsigned long long>() + which * N_COUNTERS, ab.counters, N_COUNTERS * sizeof(unsigned long long),
                            cudaMemcpyDeviceToHost, st));
  return CATT_OK;
}

int front_queue(catt_ctx* c, const ReadsDev& rv, int64_t r0, int* launches) {
  int rc;
  if ((rc = fork_j(c))) return rc;
  for (int which = 0; which < 2; which++) {
    RefSet& rs = which ? c->J : c->V;
    if ((rc = ensure_align_buffers(c->ab[which], rs.dev, rv.n)) || (!c->ab[which].tasks && (rc = ensure_task_buffers(c->ab[which], rv.n * 10 + 4096)))) return rc;
    if ((rc = front_launch(c, which, rv, r0, true, launches))) return rc;
  }
  return CATT_OK;
}

// after join_streams: the task buffer may have been too small (the piece's expansion + SW + pre-selection is re-run once), counters
int front_finish(catt_ctx* c, const ReadsDev& rv, int64_t r0, catt_info* info, int* launches) {
  int rc;
  for (int which = 0; which < 2; which++) {
    AlignBuffers& ab = c->ab[which];
    cudaStream_t st = aln_stream(c, which);
    cudaEvent_t* ev = c->ev_a[which];
    unsigned long long* cnt = c->pin[PIN_CNT].as<unsigned long long>() + which * N_COUNTERS;
    float ms_seed = 0, ms_exp = 0, ms_sw = 0, ms_sel = 0;
    cudaEventElapsedTime(&ms_seed, ev[0], ev[1]);
    for (int attempt = 0;; attempt++) {
      if (cnt[6] > 0) { set_error("%llu reads exceed the seeding run-list capacity", cnt[6]); return CATT_E_LIMIT; }
      float a = 0, b = 0, d = 0;
      cudaEventElapsedTime(&a, ev[2], ev[3]); cudaEventElapsedTime(&b, ev[3], ev[4]); cudaEventElapsedTime(&d, ev[4], ev[5]);
      ms_exp += a; ms_sw += b; ms_sel += d;
      if (cnt[9] == 0) break;
      if (attempt == 1) { set_error("SW task buffer overflow persists (%llu tasks)", cnt[9]); return CATT_E_LIMIT; }
      if ((rc = ensure_task_buffers(ab, (int64_t)cnt[9] + (int64_t)cnt[9] / 8 + 1024))) return rc;
      unsigned long long z[N_COUNTERS] = {0};
      z[4] = cnt[4]; z[6] = cnt[6];
      CATT_CUDA(cudaMemcpyAsync(ab.counters, z, sizeof z, cudaMemcpyHostToDevice, st));
      CATT_CUDA(cudaStreamSynchronize(st));
      if ((rc = front_launch(c, which, rv, r0, false, launches))) return rc;
      CATT_CUDA(cudaStreamSynchronize(st));
    }
    if (which == 0) { info->ms_seed += ms_seed + ms_exp; info->ms_seed_kernel += ms_seed; info->ms_sw += ms_sw; info->ms_select += ms_sel; }
    else info->ms_j_stream += ms_seed + ms_exp + ms_sw + ms_sel;
    if (which) { info->cand_j += (int64_t)cnt[0]; info->cells_j += (int64_t)cnt[1]; info->cells_computed_j += (int64_t)cnt[8]; }
    else { info->cand_v += (int64_t)cnt[0]; info->cells_v += (int64_t)cnt[1]; info->cells_computed_v += (int64_t)cnt[8]; }
  }
  return CATT_OK;
}

// the back half of one piece: tie-break with the reads' global ordinals, TRACK pass on the winners, CIGAR facts, traceback
int back_chunk(catt_ctx* c, const ReadsDev& rv, int64_t r0, int64_t ordinal, catt_info* info, int* launches) {
  int rc;
  if ((rc = fork_j(c))) return rc;
  for (int which = 0; which < 2; which++) {
    RefSet& rs = which ? c->J : c->V;
    AlignBuffers& ab = c->ab[which];
    cudaStream_t st = aln_stream(c, which);
    if ((rc = ensure_align_buffers(ab, rs.dev, rv.n))) return rc;
    CATT_CUDA(cudaMemsetAsync(ab.counters, 0, N_COUNTERS * sizeof(unsigned long long), st));
    CATT_CUDA(cudaEventRecord(c->ev_a[which][4], st));
    if ((rc = launch_select_post(rs.dev, rv, ab, c->df_tied[which].as<uint32_t>() + r0 * rs.dev.CW, c->df_mx[which].as<int16_t>() + r0,
                                 c->df_nc[which].as<int16_t>() + r0, c->recs_all[which].as<catt_aln>() + r0, ordinal, c->min_score, c->sm_count, st, launches))) return rc;
    CATT_CUDA(cudaEventRecord(c->ev_a[which][5], st));
    CATT_CUDA(cudaMemcpyAsync(c->pin[PIN_CNT].as<unsigned long long>() + which * N_COUNTERS, ab.counters, N_COUNTERS * sizeof(unsigned long long),
                              cudaMemcpyDeviceToHost, st));
  }
  if ((rc = join_streams(c))) return rc;
  for (int which = 0; which < 2; which++) {
    const unsigned long long* cnt = c->pin[PIN_CNT].as<unsigned long long>() + which * N_COUNTERS;
    float d = 0;
    cudaEventElapsedTime(&d, c->ev_a[which][4], c->ev_a[which][5]);
    if (which) { info->gapped_j += (int64_t)cnt[2]; info->j_hits += (int64_t)cnt[3]; info->ms_j_stream += d; }
    else { info->gapped_v += (int64_t)cnt[2]; info->v_hits += (int64_t)cnt[3]; info->ms_select += d; }
  }
  return CATT_OK;
}

int run_file_rank(catt_ctx* c, Comm* comm, int root, const char* path, catt_result** out) {
  std::unique_ptr<catt_result> R(new catt_result());
  Trace tr;
  const int N = comm->nranks, me = comm->rank;
  int launches = 0, rc = CATT_OK, status = 0;                    // status: 0 ok, 1 fallback, < 0 error
  int64_t n_loc = 0; uint64_t words = 0; int max_len = 0;
  c->shard.valid = false;
  const int fd = open(path, O_RDONLY);
  int64_t fsize = 0, b0 = 0, b1 = 0;
  if (fd < 0) { set_error("cannot read %s", path); status = CATT_E_IO; }
  if (!status) {
    fsize = (int64_t)lseek(fd, 0, SEEK_END);
    unsigned char magic[2] = {0, 0};
    if (fsize < 2 || pread(fd, magic, 2, 0) != 2 || magic[0] != '@') status = RANK_FALLBACK;         // gzip / BGZF / FASTA / empty: host path
  }
  if (!status) {
    b0 = record_boundary(fd, fsize, fsize / N * me);
    b1 = me + 1 == N ? fsize : record_boundary(fd, fsize, fsize / N * (me + 1));
    if (b0 < 0 || b1 < 0) status = RANK_FALLBACK;
  }
  // The block goes through in record-aligned pieces: piece k is read (all host threads of this rank) into one of two pinned buffers
  // and copied to HBM; the GPU indexes + packs it and runs the FRONT half of Stage A on its reads - seeding, SW, and the selection up
  // to the tie-break (select_kernel MODE 1) - while the host reads piece k+1.  The BACK half (tie-break with the global ordinal,
  // TRACK pass, CIGAR facts) follows once every rank has counted its reads.
  const double t0 = now_ms();
  struct Piece { int64_t r0, n; };
  std::vector<Piece> pieces;
  catt_info& I = R->info;
  if (!status && b1 > b0) {
    const int64_t bytes = b1 - b0;
    int64_t piece = (int64_t)64 << 20;
    if (const char* e = getenv("CATT_GROUP_PIECE")) piece = std::max<int64_t>(atoll(e), 4096);      // tests: many pieces on a small file
    if ((rc = c->rd_text.ensure((size_t)bytes + 32)) || (rc = c->pin[PIN_PACK0].ensure((size_t)piece + 16)) || (rc = c->pin[PIN_PACK1].ensure((size_t)piece + 16))) status = rc;
    char* d_text = c->rd_text.as<char>();
    int64_t file_pos = b0, text_pos = 0, tail = 0;                // next byte to read; device offset of the next piece; bytes carried over
    int64_t pending_r0 = -1, pending_n = 0;                       // the piece whose front half is in flight
    auto finish_pending = [&]() -> int {
      if (pending_r0 < 0) return CATT_OK;
      int r = join_streams(c);
      if (r) return r;
      DeviceReads dr;
      dr.words = c->rd_words.as<uint32_t>(); dr.off = c->rd_off.as<uint32_t>(); dr.len = c->rd_len.as<int32_t>(); dr.n = pending_r0 + pending_n; dr.max_len = max_len;
      r = front_finish(c, dr.view(pending_r0, pending_n), pending_r0, &I, &launches);
      pending_r0 = -1;
      return r;
    };
    for (int64_t k = 0; !status && (file_pos < b1 || tail > 0); k++) {
      char* h = c->pin[PIN_PACK0 + (k & 1)].as<char>();
      // (the carried tail was written into this buffer after its previous copy had completed, see below)
      const int64_t want = std::min<int64_t>(piece - tail, b1 - file_pos);
      const int64_t sub = 4 << 20, ns = (want + sub - 1) / sub;
      bool ok = true;
      #pragma omp parallel for schedule(dynamic, 1) reduction(&& : ok)
      for (int64_t q = 0; q < ns; q++) {
        int64_t a = q * sub;
        const int64_t e = std::min(want, a + sub);
        while (a < e) { const ssize_t r = pread(fd, h + tail + a, (size_t)(e - a), (off_t)(file_pos + a)); if (r <= 0) { ok = false; break; } a += r; }
      }
      if (!ok) { set_error("short read on %s", path); status = CATT_E_IO; break; }
      int64_t have = tail + want;
      file_pos += want;
      int64_t cut = have;
      if (file_pos >= b1) {                                      // last piece of the block: trailing blank space off, exactly one '\n'
        while (cut > 0 && (h[cut - 1] == '\n' || h[cut - 1] == '\r' || h[cut - 1] == ' ' || h[cut - 1] == '\t')) cut--;
        if (cut == 0) { if (text_pos == 0) break; status = RANK_FALLBACK; break; }     // (only blank lines left: let the host parser decide)
        h[cut++] = '\n';
        have = cut;
      } else {                                                   // the last record start in the buffer: a line '@...' whose second next line is '+...'
        int64_t q = have;
        cut = -1;
        for (int tries = 0; tries < 64 && q > 0; tries++) {
          const char* nl = (const char*)memrchr(h, '\n', (size_t)(q - 1));      // the newline before the line that ends at or before q
          const int64_t ls = nl ? (nl - h) + 1 : 0;
          if (h[ls] == '@' && ls > 0) {
            const char* n1 = (const char*)memchr(h + ls, '\n', (size_t)(have - ls));
            const char* n2 = n1 ? (const char*)memchr(n1 + 1, '\n', (size_t)(h + have - (n1 + 1))) : nullptr;
            if (n2 && n2 + 1 < h + have && n2[1] == '+') { cut = ls; break; }
          }
          q = ls;
        }
        if (cut <= 0) { status = RANK_FALLBACK; break; }          // no record boundary in a whole piece: not the FASTQ this path handles
      }
      if (cudaMemcpyAsync(d_text + text_pos, h, (size_t)cut, cudaMemcpyHostToDevice, c->copy_stream) != cudaSuccess ||
          cudaEventRecord(c->ev_h2d[k & 1], c->copy_stream) != cudaSuccess) { set_error("FASTQ text upload failed"); status = CATT_E_CUDA; break; }
      I.bytes_h2d += cut;
      // the unfinished record(s) behind the cut go to the front of the other buffer, once the copy that last read it is done
      tail = have - cut;
      if (tail > 0) {
        if (k >= 1 && cudaEventSynchronize(c->ev_h2d[(k + 1) & 1]) != cudaSuccess) { set_error("FASTQ text upload failed"); status = CATT_E_CUDA; break; }
        memcpy(c->pin[PIN_PACK0 + ((k + 1) & 1)].as<char>(), h + cut, (size_t)tail);
      } else if (k >= 1 && cudaEventSynchronize(c->ev_h2d[(k + 1) & 1]) != cudaSuccess) { set_error("FASTQ text upload failed"); status = CATT_E_CUDA; break; }
      // the previous piece's kernels must be done before its per-piece scratch is reused (they ran while this piece was read)
      if ((rc = finish_pending())) { status = rc; break; }
      // index + pack this piece, then queue the front half of Stage A on its reads
      if (cudaStreamWaitEvent(c->stream, c->ev_h2d[k & 1], 0) != cudaSuccess) { set_error("FASTQ text upload failed"); status = CATT_E_CUDA; break; }
      int fb = 0;
      int64_t np = 0; uint64_t wp = 0;
      rc = ingest_fastq(c, d_text, text_pos, text_pos + cut, n_loc, words, &np, &wp, &max_len, &fb, &launches, c->stream);
      if (rc) { status = fb ? RANK_FALLBACK : rc; break; }
      if (k == 0 && np > 0 && bytes > 2 * cut) {                  // size the per-read arrays for the whole block once (each growth is a copy)
        const int64_t est = (int64_t)((double)bytes / ((double)cut / (double)np) * 1.03) + 4096;
        const uint64_t est_words = (uint64_t)((double)wp / (double)np * (double)est) + 4096;
        cudaStream_t st = c->stream;
        if (est_words < 0xFFFFFFF0ull &&
            ((rc = c->rd_name_beg.ensure_keep((size_t)(est + 1) * 8, st)) || (rc = c->rd_name_end.ensure_keep((size_t)(est + 1) * 8, st)) ||
             (rc = c->rd_seq_beg.ensure_keep((size_t)(est + 1) * 8, st)) || (rc = c->rd_qual_beg.ensure_keep((size_t)(est + 1) * 8, st)) ||
             (rc = c->rd_len.ensure_keep((size_t)(est + 1) * 4, st)) || (rc = c->rd_off.ensure_keep((size_t)(est + 2) * 4, st)) ||
             (rc = c->rd_words.ensure_keep((size_t)(est_words + 4) * 4, st)) || (rc = deferred_reserve(c, est, st)))) { status = rc; break; }
      }
      if ((rc = deferred_reserve(c, n_loc + np, c->stream))) { status = rc; break; }
      if (np > 0) {
        DeviceReads dr;
        dr.words = c->rd_words.as<uint32_t>(); dr.off = c->rd_off.as<uint32_t>(); dr.len = c->rd_len.as<int32_t>(); dr.n = n_loc + np; dr.max_len = max_len;
        if ((rc = front_queue(c, dr.view(n_loc, np), n_loc, &launches))) { status = rc; break; }
        pending_r0 = n_loc; pending_n = np;
        pieces.push_back(Piece{n_loc, np});
      }
      n_loc += np; words += wp; text_pos += cut;
    }
    if (!status && (rc = finish_pending())) status = rc;
    if (status) { join_streams(c); cudaStreamSynchronize(c->copy_stream); }
    tr.lap("run_file_rank: read + H2D + index + pack + seeding / SW, piece by piece");
  }
  if (fd >= 0) close(fd);
  const double ms_ingest = now_ms() - t0;
  // first exchange: read counts and status.  Every rank answers it, whatever happened.
  std::string err = status < 0 ? catt::last_error() : "";
  long long mine[2] = {n_loc, status};
  std::vector<long long> all((size_t)2 * N);
  if ((rc = comm->host_allgather(mine, all.data(), sizeof mine, c->stream))) return rc;
  int64_t lo = 0, n_total = 0;
  bool fallback = false;
  for (int r = 0; r < N; r++) {
    if (all[2 * r + 1] < 0) { if (r == me) set_error("%s", err.c_str()); else set_error("GPU group: rank %d could not ingest %s (code %lld)", r, path, all[2 * r + 1]); return (int)all[2 * r + 1]; }
    fallback = fallback || all[2 * r + 1] == RANK_FALLBACK;
    if (r < me) lo += all[2 * r];
    n_total += all[2 * r];
  }
  if (fallback) return RANK_FALLBACK;
  c->last_n = n_loc; c->last_words = (int64_t)words; c->last_max_len = max_len;
  I.ms_pack = ms_ingest;
  // the back half, now that this block's first read has its global ordinal
  const double ta = now_ms();
  rc = ensure_recs(c, n_loc);
  FileHits fh;
  fh.start = {0, n_loc};
  fh.v.assign(1, 0); fh.j.assign(1, 0);
  for (size_t k = 0; !rc && k < pieces.size(); k++) {
    DeviceReads dr;
    dr.words = c->rd_words.as<uint32_t>(); dr.off = c->rd_off.as<uint32_t>(); dr.len = c->rd_len.as<int32_t>(); dr.n = n_loc; dr.max_len = max_len;
    rc = back_chunk(c, dr.view(pieces[k].r0, pieces[k].n), pieces[k].r0, lo + pieces[k].r0, &I, &launches);
  }
  const double ms_a = now_ms() - ta + ms_ingest;
  tr.lap("run_file_rank: tie-break + TRACK pass");
  NamesDev nd;
  nd.s = c->rd_text.as<char>(); nd.beg = reinterpret_cast<const int64_t*>(c->rd_name_beg.p); nd.end = reinterpret_cast<const int64_t*>(c->rd_name_end.p);
  DeviceQuals dq;
  dq.quals = c->rd_text.as<char>(); dq.seq_off = reinterpret_cast<const int64_t*>(c->rd_qual_beg.p);
  return rank_finish(c, comm, root, n_total, lo, n_loc, {0, n_total}, nd, dq, max_len, ms_a, rc, R, launches, out);
}

int run_reads_dispatch(catt_ctx* c, const HostReads& hr, const std::vector<int64_t>& fs, catt_result** out);
// Runs `body(r, rank context, communicator)` on every rank of the in-process group: rank r on its own host thread and device, rank 0
// on the calling thread.  Returns the first real error (a peer's "rank failed" echo is not one), or `same` when every rank returned
// the same non-error code.
int group_run(catt_ctx* c, const std::function<int(int, catt_ctx*, Comm*, catt_result**)>& body, catt_result** out) {
  const int N = 1 + (int)c->peers.size();
  std::vector<int> rcs((size_t)N, CATT_OK);
  std::vector<std::string> errs((size_t)N);
  std::vector<catt_result*> res((size_t)N, nullptr);
  const int host_threads = std::max(1, omp_get_max_threads()), per_rank = std::max(1, host_threads / N);
  if (c->lgroup) catt::local_group_reset(c->lgroup);
  auto run = [&](int r) {
    catt_ctx* cr = r ? c->peers[(size_t)r - 1] : c;
    if (cudaSetDevice(cr->device) != cudaSuccess) { rcs[r] = CATT_E_CUDA; errs[r] = "cudaSetDevice failed"; return; }
    omp_set_num_threads(per_rank);                   // a per-thread setting: every rank packs / copies with its share of the cores
    rcs[r] = body(r, cr, c->comms[(size_t)r], &res[r]);
    if (rcs[r] < 0) { errs[r] = catt::last_error(); if (c->lgroup) catt::local_group_fail(c->lgroup); }
  };
  std::vector<std::thread> th;
  for (int r = 1; r < N; r++) th.emplace_back(run, r);
  run(0);
  for (auto& t : th) t.join();
  omp_set_num_threads(host_threads);
  cudaSetDevice(c->device);
  int bad = -1;
  for (int r = 0; r < N; r++) if (rcs[r] < 0 && (bad < 0 || errs[bad].find("peer rank") != std::string::npos)) bad = r;     // prefer the root cause
  if (bad >= 0) {
    for (auto* p : res) delete p;
    set_error("GPU group rank %d (device %d): %s", bad, bad ? c->peers[(size_t)bad - 1]->device : c->device, errs[bad].c_str());
    return rcs[bad];
  }
  for (int r = 1; r < N; r++) delete res[r];
  if (rcs[0] != CATT_OK) { delete res[0]; return rcs[0]; }
  *out = res[0];
  return CATT_OK;
}

int run_reads_dispatch(catt_ctx* c, const HostReads& hr, const std::vector<int64_t>& fs, catt_result** out);
// File input on a GPU group.  One plain FASTQ file: every rank ingests its byte range on its own GPU (run_file_rank).  Anything
// else (several files, .gz / BGZF, BAM / SAM, FASTA, wrapped lines): read and parsed on the host, then split over the ranks.
int run_files_group(catt_ctx* c, int nfiles, const char* const* paths, bool bam, catt_result** out) {
  if (nfiles == 1 && !bam && paths[0] && !getenv("CATT_HOST_PARSE")) {
    const char* path = paths[0];
    int rc = group_run(c, [&](int, catt_ctx* cr, Comm* comm, catt_result** o) { return run_file_rank(cr, comm, 0, path, o); }, out);
    if (rc != RANK_FALLBACK) return rc;
  }
  HostReads hr;
  std::vector<int64_t> fs{0};
  Trace tr;
  for (int f = 0; f < nfiles; f++) {
    if (!paths[f]) { set_error("catt_run_files: null path"); return CATT_E_ARG; }
    int rc = bam ? parse_bam(paths[f], &hr) : parse_fastx(paths[f], &hr);
    if (rc) return rc;
    fs.push_back(hr.n);
  }
  tr.lap("run_files_group: read + parse");
  return run_reads_dispatch(c, hr, fs, out);
}

int run_reads_dispatch(catt_ctx* c, const HostReads& hr, const std::vector<int64_t>& fs, catt_result** out) {
  if (c->peers.empty()) return run_all(c, hr, fs, out);
  const int N = 1 + (int)c->peers.size();
  return group_run(c, [&](int r, catt_ctx* cr, Comm* comm, catt_result** o) {
    int64_t lo, hi;
    block_bounds(hr.n, N, r, &lo, &hi);
    HostReads sub;
    sub.n = hi - lo; sub.names = hr.names; sub.name_off = hr.name_off ? hr.name_off + lo : nullptr; sub.seqs = hr.seqs;
    sub.seq_off = hr.seq_off ? hr.seq_off + lo : nullptr; sub.quals = hr.quals;
    return run_rank(cr, comm, 0, hr.n, lo, sub, fs, o);
  }, out);
}

std::string dup(const char* s) { return s ? std::string(s) : std::string(); }

inline uint64_t sm64(uint64_t& s) {
  s += 0x9E3779B97F4A7C15ull;
  uint64_t z = s;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

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
  // a group of GPUs for one sample: this context is rank 0 (devices[0]); the other ranks get contexts of their own below
  const int ngpu = cfg->ngpu > 1 ? cfg->ngpu : 1;
  if (ngpu > 16) { set_error("ngpu = %d exceeds the limit of 16 ranks", ngpu); return CATT_E_ARG; }
  std::vector<int> devs((size_t)ngpu);
  for (int r = 0; r < ngpu; r++) devs[r] = ngpu > 1 ? (cfg->devices ? cfg->devices[r] : r) : cfg->device;
  for (int r = 0; r < ngpu; r++)
    if (devs[r] < 0 || devs[r] >= ndev) { set_error("device %d out of range (have %d)", devs[r], ndev); return CATT_E_ARG; }
  if (cfg->kmer != CATT_KMER_AUTO && (cfg->kmer < 1 || cfg->kmer > KMER_MAX)) {
    // args["kmer"] is free in the reference's CLI (catt:18); the k-mer table here holds one 2-bit packed 64-bit word per k-mer
    set_error("kmer = %d is outside the supported range 1..%d (or CATT_KMER_AUTO)", cfg->kmer, KMER_MAX);
    return CATT_E_LIMIT;
  }
  const int device = devs[0];
  CATT_CUDA(cudaSetDevice(device));
  // (every early return below releases what was allocated so far: device tables, streams, events, the other ranks' contexts)
  std::unique_ptr<catt_ctx, void (*)(catt_ctx*)> c(new catt_ctx(), catt_ctx_destroy);
  c->device = device;
  cudaDeviceProp prop;
  CATT_CUDA(cudaGetDeviceProperties(&prop, device));
  c->sm_count = prop.multiProcessorCount;
  c->min_score = cfg->min_score; c->kmer = cfg->kmer; c->nthreads = cfg->julia_nthreads < 1 ? 1 : cfg->julia_nthreads;
  c->allow_v = cfg->allow_v; c->allow_j = cfg->allow_j;
  int rc;
  if ((rc = c->V.load(cfg->v_fasta)) || (rc = c->J.load(cfg->j_fasta))) return rc;
  if ((rc = c->V.upload()) || (rc = c->J.upload())) return rc;
  // J sets (9-62 short records): a read has one or two J candidates, so each is scored once by the TRACK kernel (sw.cu: launch_sw)
  c->J.dev.direct = getenv("CATT_NO_DIRECT") ? 0 : 1;
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
  c->keep_records = cfg->keep_records != 0;
  c->only_cdr3 = cfg->only_cdr3 != 0;
  c->want_clonotypes = cfg->clonotypes != 0;
  if (cfg->chunk_reads > 0) c->chunk_reads = cfg->chunk_reads;
  if (const char* e = getenv("CATT_CHUNK_READS")) { const long long v = atoll(e); if (v > 0) c->chunk_reads = v; }
  CATT_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CATT_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  CATT_CUDA(cudaStreamCreateWithFlags(&c->stream_j, cudaStreamNonBlocking));
  CATT_CUDA(cudaStreamCreateWithFlags(&c->stream_ingest, cudaStreamNonBlocking));
  CATT_CUDA(cudaEventCreateWithFlags(&c->ev_ingest, cudaEventDisableTiming));
  CATT_CUDA(cudaEventCreateWithFlags(&c->ev_inputs, cudaEventDisableTiming));
  for (auto& ev : c->ev) CATT_CUDA(cudaEventCreate(&ev));
  for (auto& row : c->ev_a) for (auto& ev : row) CATT_CUDA(cudaEventCreate(&ev));
  for (auto& ev : c->ev_h2d) CATT_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  for (auto& ev : c->ev_q) CATT_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  if ((rc = c->pin[PIN_CNT].ensure(1024))) return rc;
  if (ngpu > 1) {
    struct Guard { catt_ctx* c; bool armed = true; ~Guard() { if (armed) { for (auto* p : c->peers) catt_ctx_destroy(p); c->peers.clear(); for (auto* m : c->comms) delete m; c->comms.clear(); } } } guard{c.get()};
    for (int r = 1; r < ngpu; r++) {
      catt_config sub = *cfg;
      sub.ngpu = 1; sub.devices = nullptr; sub.device = devs[r];
      catt_ctx* p = nullptr;
      if ((rc = catt_ctx_create(&p, &sub))) return rc;
      c->peers.push_back(p);
    }
    CATT_CUDA(cudaSetDevice(device));
    bool distinct = true;
    for (int a = 0; a < ngpu; a++) for (int b = a + 1; b < ngpu; b++) distinct = distinct && devs[a] != devs[b];
    const char* force = getenv("CATT_COMM");
    const bool want_local = !distinct || (force && !strcmp(force, "local"));
    if (!want_local) {
      if ((rc = nccl_comms_create_all(ngpu, devs.data(), &c->comms))) return rc;
      CATT_CUDA(cudaSetDevice(device));
    } else {
      c->lgroup = local_group_create(ngpu, devs.data());
      if (!c->lgroup) { set_error("cannot create the in-process GPU group"); return CATT_E_ARG; }
      for (int r = 0; r < ngpu; r++) c->comms.push_back(local_comm_create(c->lgroup, r));
    }
    guard.armed = false;
  }
  *out = c.release();
  return CATT_OK;
}

void catt_ctx_destroy(catt_ctx* c) {
  if (!c) return;
  for (size_t r = 0; r < c->comms.size(); r++) {       // communicators first (each on its own device), then the other ranks' contexts
    cudaSetDevice(r ? c->peers[r - 1]->device : c->device);
    delete c->comms[r];
  }
  c->comms.clear();
  for (auto* p : c->peers) catt_ctx_destroy(p);
  c->peers.clear();
  cudaSetDevice(c->device);
  delete c->comm;
  for (auto& b : c->sh) b.release();
  for (auto& b : c->cl) b.release();
  for (int w = 0; w < 2; w++) { c->df_tied[w].release(); c->df_mx[w].release(); c->df_nc[w].release(); }
  c->V.release(); c->J.release();
  free_align_buffers(c->ab[0]); free_align_buffers(c->ab[1]);
  for (auto& b : c->sb) b.release();
  for (auto& b : c->pin) b.release();
  for (auto& b : c->recs_all) b.release();
  c->rd_words.release(); c->rd_off.release(); c->rd_len.release(); c->rd_quals.release(); c->rd_seqoff.release();
  c->rd_text.release(); c->rd_nl.release(); c->rd_name_beg.release(); c->rd_name_end.release(); c->rd_seq_beg.release();
  c->rd_qual_beg.release(); c->rd_nwords.release();
  for (auto& b : c->nb) b.release();
  cudaFree(c->d_finder); cudaFree(c->d_dseq); cudaFree(c->d_doff);
  for (auto& ev : c->ev) if (ev) cudaEventDestroy(ev);
  for (auto& row : c->ev_a) for (auto& ev : row) if (ev) cudaEventDestroy(ev);
  for (auto& ev : c->ev_h2d) if (ev) cudaEventDestroy(ev);
  for (auto& ev : c->ev_q) if (ev) cudaEventDestroy(ev);
  for (auto& ev : c->ev_piece) if (ev) cudaEventDestroy(ev);
  for (auto& b : c->rd_raw) b.release();
  if (c->stream) cudaStreamDestroy(c->stream);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->stream_j) cudaStreamDestroy(c->stream_j);
  if (c->stream_ingest) cudaStreamDestroy(c->stream_ingest);
  if (c->ev_ingest) cudaEventDestroy(c->ev_ingest);
  if (c->ev_inputs) cudaEventDestroy(c->ev_inputs);
  delete c;
}

int catt_set_host_threads(int32_t n) {
  if (n < 0) { set_error("catt_set_host_threads: negative thread count"); return CATT_E_ARG; }
  omp_set_num_threads(n > 0 ? n : omp_get_num_procs());
  return CATT_OK;
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
  return run_reads_dispatch(c, hr, {0, n}, out);
}

int catt_run_reads_files(catt_ctx* c, int64_t n, const char* names, const int64_t* name_off, const char* seqs, const int64_t* seq_off,
                         const char* quals, int32_t nfiles, const int64_t* file_start, catt_result** out) {
  if (!c || !out || n < 0 || nfiles < 1 || !file_start || (n > 0 && (!names || !name_off || !seqs || !seq_off))) {
    set_error("catt_run_reads_files: bad argument");
    return CATT_E_ARG;
  }
  std::vector<int64_t> fs(file_start, file_start + nfiles + 1);
  if (fs.front() != 0 || fs.back() != n || !std::is_sorted(fs.begin(), fs.end())) { set_error("catt_run_reads_files: file_start must run from 0 to n"); return CATT_E_ARG; }
  CATT_CUDA(cudaSetDevice(c->device));
  HostReads hr;
  hr.n = n; hr.names = names; hr.name_off = name_off; hr.seqs = seqs; hr.seq_off = seq_off; hr.quals = quals;
  return run_reads_dispatch(c, hr, fs, out);
}

int catt_run_file(catt_ctx* c, const char* path, catt_result** out) {
  if (!c || !path || !out) { set_error("catt_run_file: null argument"); return CATT_E_ARG; }
  const char* paths[1] = {path};
  return catt_run_files(c, 1, paths, out);
}

int catt_run_files(catt_ctx* c, int32_t nfiles, const char* const* paths, catt_result** out) {
  if (!c || !paths || !out || nfiles < 1) { set_error("catt_run_files: bad argument"); return CATT_E_ARG; }
  CATT_CUDA(cudaSetDevice(c->device));
  if (!c->peers.empty()) return run_files_group(c, nfiles, paths, false, out);
  if (!getenv("CATT_HOST_PARSE")) {
    bool handled = false;
    int rc = run_files_device(c, nfiles, paths, out, &handled);
    if (rc || handled) return rc;
  }
  HostReads hr;
  std::vector<int64_t> fs{0};
  Trace tr;
  for (int f = 0; f < nfiles; f++) {
    if (!paths[f]) { set_error("catt_run_files: null path"); return CATT_E_ARG; }
    int rc = parse_fastx(paths[f], &hr);
    if (rc) return rc;
    fs.push_back(hr.n);
  }
  tr.lap("run_files: read + parse");
  return run_reads_dispatch(c, hr, fs, out);
}

int catt_run_bam(catt_ctx* c, const char* path, catt_result** out) {
  if (!c || !path || !out) { set_error("catt_run_bam: null argument"); return CATT_E_ARG; }
  CATT_CUDA(cudaSetDevice(c->device));
  if (!c->peers.empty()) { const char* one[1] = {path}; return run_files_group(c, 1, one, true, out); }
  if (!getenv("CATT_HOST_PARSE")) {                      // BGZF BAM: converted to FASTQ text group by group under the GPU pipeline
    bool handled = false;
    const char* paths[1] = {path};
    int rc = run_files_device(c, 1, paths, out, &handled, true);
    if (rc || handled) return rc;
  }
  HostReads hr;
  Trace tr;
  int rc = parse_bam(path, &hr);
  if (rc) return rc;
  tr.lap("run_bam: inflate + decode");
  return run_reads_dispatch(c, hr, {0, hr.n}, out);
}

int catt_inflate_file(const char* path, const char* out_path, int32_t* method) {
  if (!path || !out_path) { set_error("catt_inflate_file: null argument"); return CATT_E_ARG; }
  std::string txt;
  int how = 0;
  if (!slurp_file(path, txt, &how)) { set_error("cannot read %s", path); return CATT_E_IO; }
  FILE* f = fopen(out_path, "wb");
  if (!f) { set_error("cannot write %s", out_path); return CATT_E_IO; }
  const bool ok = fwrite(txt.data(), 1, txt.size(), f) == txt.size();
  if (fclose(f) != 0 || !ok) { set_error("short write to %s", out_path); return CATT_E_IO; }
  if (method) *method = how;
  return CATT_OK;
}

int catt_bam_to_fastq(const char* bam_path, const char* fastq_path) {
  if (!bam_path || !fastq_path) { set_error("catt_bam_to_fastq: null argument"); return CATT_E_ARG; }
  if (!getenv("CATT_HOST_PARSE")) {                      // BGZF BAM: the streamed converter the file pipeline uses
    RawFile raw;
    BamTextStream bt;
    if (raw.open(bam_path) && raw.size() >= 2 && raw.data()[0] == 0x1f && raw.data()[1] == 0x8b && bt.open(raw.data(), raw.size())) {
      FILE* f = fopen(fastq_path, "wb");
      if (!f) { set_error("cannot write %s", fastq_path); return CATT_E_IO; }
      std::vector<char> text((size_t)std::max(1, omp_get_max_threads()) * 48 * 65536 / 3 * 4 + (1u << 20));
      int r = 1;
      bool wrote_ok = true;
      while (r == 1) {
        size_t wrote = 0;
        r = bt.next(text.data(), text.size(), &wrote);
        if (r >= 0 && wrote && fwrite(text.data(), 1, wrote, f) != wrote) wrote_ok = false;
      }
      const bool closed = fclose(f) == 0;
      if (r == 0) {
        if (!wrote_ok || !closed) { set_error("short write to %s", fastq_path); return CATT_E_IO; }
        return CATT_OK;
      }
      // malformed for the streamed converter: the host decoder below names the problem
    }
  }
  HostReads hr;
  int rc = parse_bam(bam_path, &hr);
  if (rc) return rc;
  FILE* f = fopen(fastq_path, "wb");
  if (!f) { set_error("cannot write %s", fastq_path); return CATT_E_IO; }
  std::string buf;
  for (int64_t i = 0; i < hr.n; i++) {
    const int64_t s0 = hr.seq_off[i], s1 = hr.seq_off[i + 1];
    buf += '@'; buf.append(hr.names + hr.name_off[i], (size_t)(hr.name_off[i + 1] - hr.name_off[i])); buf += '\n';
    buf.append(hr.seqs + s0, (size_t)(s1 - s0)); buf += "\n+\n";
    buf.append(hr.quals + s0, (size_t)(s1 - s0)); buf += '\n';
    if (buf.size() > (1u << 24) || i + 1 == hr.n) {
      if (fwrite(buf.data(), 1, buf.size(), f) != buf.size()) { fclose(f); set_error("short write to %s", fastq_path); return CATT_E_IO; }
      buf.clear();
    }
  }
  if (fclose(f) != 0) { set_error("cannot close %s", fastq_path); return CATT_E_IO; }
  return CATT_OK;
}

int catt_comm_id(uint8_t id[128]) {
  if (!id) { set_error("catt_comm_id: null argument"); return CATT_E_ARG; }
  return nccl_unique_id(id);
}

int catt_ctx_join(catt_ctx* c, int32_t nranks, int32_t rank, const uint8_t id[128]) {
  if (!c || !id) { set_error("catt_ctx_join: null argument"); return CATT_E_ARG; }
  if (!c->peers.empty() || c->comm) { set_error("catt_ctx_join: the context already belongs to a GPU group"); return CATT_E_ARG; }
  CATT_CUDA(cudaSetDevice(c->device));
  return nccl_comm_create_rank(nranks, rank, id, &c->comm);
}

int catt_run_reads_shard(catt_ctx* c, int64_t n_total, int64_t lo, int64_t n_local, const char* names, const int64_t* name_off, const char* seqs,
                         const int64_t* seq_off, const char* quals, int32_t nfiles, const int64_t* file_start, catt_result** out) {
  if (!c || !out || n_total < 0 || lo < 0 || n_local < 0 || nfiles < 1 || !file_start || (n_local > 0 && (!names || !name_off || !seqs || !seq_off))) {
    set_error("catt_run_reads_shard: bad argument");
    return CATT_E_ARG;
  }
  if (!c->comm) { set_error("catt_run_reads_shard: call catt_ctx_join first"); return CATT_E_ARG; }
  std::vector<int64_t> fs(file_start, file_start + nfiles + 1);
  if (fs.front() != 0 || fs.back() != n_total || !std::is_sorted(fs.begin(), fs.end())) { set_error("catt_run_reads_shard: file_start must run from 0 to n_total"); return CATT_E_ARG; }
  CATT_CUDA(cudaSetDevice(c->device));
  static const int64_t zero_off[1] = {0};
  HostReads hr;
  hr.n = n_local; hr.names = names; hr.name_off = n_local ? name_off : zero_off; hr.seqs = seqs; hr.seq_off = n_local ? seq_off : zero_off; hr.quals = quals;
  return run_rank(c, c->comm, 0, n_total, lo, hr, fs, out);
}

int catt_run_file_shard(catt_ctx* c, const char* path, catt_result** out) {
  if (!c || !path || !out) { set_error("catt_run_file_shard: null argument"); return CATT_E_ARG; }
  if (!c->comm) { set_error("catt_run_file_shard: call catt_ctx_join first"); return CATT_E_ARG; }
  CATT_CUDA(cudaSetDevice(c->device));
  const int rc = run_file_rank(c, c->comm, 0, path, out);
  if (rc == RANK_FALLBACK) { set_error("%s is not a plain 4-line FASTQ file: catt_run_file_shard reads nothing else (inflate / convert it first, or use catt_run_reads_shard)", path); return CATT_E_ARG; }
  return rc;
}

// Resident form of catt_run_reads_shard (bench.py's `value`): catt_shard_upload moves this rank's block to HBM - packed reads,
// QNAMEs, quality strings - and catt_shard_run, which may be called repeatedly, runs Stage A on the resident block and the global
// Stage B with every input already on the device.
int catt_shard_upload(catt_ctx* c, int64_t n_total, int64_t lo, int64_t n_local, const char* names, const int64_t* name_off, const char* seqs,
                      const int64_t* seq_off, const char* quals, int32_t nfiles, const int64_t* file_start) {
  if (!c || n_total < 0 || lo < 0 || n_local < 0 || nfiles < 1 || !file_start || (n_local > 0 && (!names || !name_off || !seqs || !seq_off))) {
    set_error("catt_shard_upload: bad argument");
    return CATT_E_ARG;
  }
  if (!c->peers.empty()) { set_error("catt_shard_upload: not available on an in-process GPU group (use one process per GPU)"); return CATT_E_ARG; }
  CATT_CUDA(cudaSetDevice(c->device));
  static const int64_t zero_off[1] = {0};
  HostReads hr;
  hr.n = n_local; hr.names = names; hr.name_off = n_local ? name_off : zero_off; hr.seqs = seqs; hr.seq_off = n_local ? seq_off : zero_off; hr.quals = quals;
  catt_ctx::Shard& S = c->shard;
  S = catt_ctx::Shard{};
  S.gstart.assign(file_start, file_start + nfiles + 1);
  int rc;
  if ((rc = rank_files(n_total, lo, n_local, S.gstart, &S.fh, &S.ord0))) return rc;
  catt_info info{};
  int launches = 0;
  FileHits fh = S.fh;
  c->resident_quals = true;
  rc = stage_a_stream(c, hr, 0, false, true, &S.dr, &S.dq, &fh, &info, &launches);      // pack + upload only
  c->resident_quals = false;
  if (rc) return rc;
  if ((rc = upload_names(c, n_local, hr.names, hr.name_off, &S.nd, &info))) return rc;
  CATT_CUDA(cudaStreamSynchronize(c->stream));
  S.n_total = n_total; S.lo = lo; S.valid = true;
  return CATT_OK;
}

int catt_shard_run(catt_ctx* c, catt_result** out) {
  if (!c || !out) { set_error("catt_shard_run: null argument"); return CATT_E_ARG; }
  catt_ctx::Shard& S = c->shard;
  if (!S.valid) { set_error("catt_shard_run: call catt_shard_upload first"); return CATT_E_ARG; }
  CATT_CUDA(cudaSetDevice(c->device));
  std::unique_ptr<catt_result> R(new catt_result());
  int launches = 0;
  FileHits fh = S.fh;
  // (a previous run may have grown - moved - the read buffers to append the mates of straddling pairs)
  S.dr.words = c->rd_words.as<uint32_t>(); S.dr.off = c->rd_off.as<uint32_t>(); S.dr.len = c->rd_len.as<int32_t>();
  const double t0 = now_ms();
  int rc = stage_a_resident(c, S.dr, 0, &fh, &R->info, &launches, &S.ord0);
  const double ms_a = now_ms() - t0;
  if (!c->comm) {                                     // one GPU: the block is the sample
    if (rc) return rc;
    if (S.lo != 0 || S.dr.n != S.n_total) { set_error("catt_shard_run: the context is not part of a GPU group (catt_ctx_join) but holds a partial block"); return CATT_E_ARG; }
    R->info.n_reads = S.n_total;
    fh.start = S.gstart;
    if ((rc = stage_b(c, S.dr.n, S.nd, S.dr.view(), S.dq, fh, R.get(), &launches))) return rc;
    R->info.kernel_launches = launches;
    R->info.ms_stage_a_max = R->info.ms_stage_a_min = ms_a;
    *out = R.release();
    return CATT_OK;
  }
  c->last_n = S.dr.n; c->last_max_len = S.dr.max_len;   // (last_words is what the upload left)
  return rank_finish(c, c->comm, 0, S.n_total, S.lo, S.dr.n, S.gstart, S.nd, S.dq, S.dr.max_len, ms_a, rc, R, launches, out);
}

// Order-sensitive digest of both result lists (every field of every catt_read and both strings): equal digests = identical lists.
int catt_result_digest(const catt_result* r, uint64_t out[2]) {
  if (!r || !out) return CATT_E_ARG;
  for (int which = 0; which < 2; which++) {
    const catt_read* rd = r->part[which];
    const int64_t n = r->npart[which];
    uint64_t acc = 0;
    #pragma omp parallel for reduction(+ : acc) schedule(static)
    for (int64_t i = 0; i < n; i++) {
      uint64_t h = 0xcbf29ce484222325ull ^ (uint64_t)i * 0x9E3779B97F4A7C15ull;
      auto mix = [&](const void* p, size_t len) { const unsigned char* b = (const unsigned char*)p; for (size_t k = 0; k < len; k++) { h ^= b[k]; h *= 0x100000001b3ull; } };
      mix(rd[i].seq, (size_t)rd[i].seq_len); mix(rd[i].qual, (size_t)rd[i].qual_len);
      mix(&rd[i].seq_len, 4 * 7); mix(&rd[i].able, 2);
      h ^= h >> 29; h *= 0xbf58476d1ce4e5b9ull; h ^= h >> 32;
      acc += h;
    }
    out[which] = acc ^ ((uint64_t)n * 0x94d049bb133111ebull);
  }
  return CATT_OK;
}

// Canonical digest of a list as TEXT (what Julia would see): per entry FNV-1a over seq 0x1f qual 0x1f V name 0x1f J name 0x1f cdr3
// 0x1f able, mixed with the entry's position, summed.  Independent of ids and offsets, so any implementation of the path can
// compute it (the CPU checker under tests does) and two runs can be compared without moving the lists.
int catt_result_text_digest(const catt_ctx* c, const catt_result* r, uint64_t out[2]) {
  if (!c || !r || !out) return CATT_E_ARG;
  for (int which = 0; which < 2; which++) {
    const catt_read* rd = r->part[which];
    const int64_t n = r->npart[which];
    uint64_t acc = 0;
    #pragma omp parallel for reduction(+ : acc) schedule(static)
    for (int64_t i = 0; i < n; i++) {
      uint64_t h = 0xcbf29ce484222325ull ^ ((uint64_t)i * 0x9E3779B97F4A7C15ull);
      auto mix = [&](const char* p, size_t len) { for (size_t k = 0; k < len; k++) { h ^= (unsigned char)p[k]; h *= 0x100000001b3ull; } h ^= 0x1f; h *= 0x100000001b3ull; };
      const catt_read& e = rd[i];
      mix(e.seq, (size_t)e.seq_len); mix(e.qual, (size_t)e.qual_len);
      const std::string* vn = e.v_id >= 0 && e.v_id < (int)c->V.names.size() ? &c->V.names[e.v_id] : nullptr;
      const std::string* jn = e.j_id >= 0 && e.j_id < (int)c->J.names.size() ? &c->J.names[e.j_id] : nullptr;
      if (vn) mix(vn->data(), vn->size()); else mix("None", 4);
      if (jn) mix(jn->data(), jn->size()); else mix("None", 4);
      if (e.cdr3_off >= 0) mix(e.seq + e.cdr3_off, (size_t)e.cdr3_len); else mix("None", 4);
      h ^= (uint64_t)(e.able ? 1 : 0); h *= 0x100000001b3ull;
      h ^= h >> 29; h *= 0xbf58476d1ce4e5b9ull; h ^= h >> 32;
      acc += h;
    }
    out[which] = acc ^ ((uint64_t)n * 0x94d049bb133111ebull);
  }
  return CATT_OK;
}

int catt_result_info(const catt_result* r, catt_info* out) { if (!r || !out) return CATT_E_ARG; *out = r->info; return CATT_OK; }
int catt_result_reads(const catt_result* r, int which, const catt_read** reads, int64_t* n) {
  if (!r || !reads || !n || which < 0 || which > 1) return CATT_E_ARG;
  *reads = r->part[which]; *n = r->npart[which];
  return CATT_OK;
}
int catt_align_records(const catt_result* r, int which, const catt_aln** recs, int64_t* n) {
  if (!r || !recs || !n || which < 0 || which > 1) return CATT_E_ARG;
  *recs = r->rec[which].data(); *n = (int64_t)r->rec[which].size();
  return CATT_OK;
}
int catt_result_clonotypes(const catt_result* r, const catt_clonotype** rows, int64_t* n) {
  if (!r || !rows || !n) return CATT_E_ARG;
  *rows = r->clono.data(); *n = r->n_clono;
  return CATT_OK;
}
void catt_result_free(catt_result* r) { Trace tr; delete r; tr.lap("result_free"); }

int catt_best_d(catt_ctx* c, const char* const* cdr3, int64_t n, int32_t* out) {
  if (!c || !out || (n > 0 && !cdr3)) return CATT_E_ARG;
  if (c->d_seqs.empty()) { for (int64_t i = 0; i < n; i++) out[i] = -1; return CATT_OK; }
  CATT_CUDA(cudaSetDevice(c->device));
  // in chunks whose concatenated text stays far below 2^31 bytes (the offsets are int32); the context's scratch is reused
  const int64_t CH = 1 << 22;
  std::string cat; std::vector<int32_t> off;
  for (int64_t i0 = 0; i0 < n; i0 += CH) {
    const int64_t m = std::min(CH, n - i0);
    cat.clear(); off.assign(1, 0);
    for (int64_t i = 0; i < m; i++) {
      if (!cdr3[i0 + i]) { set_error("catt_best_d: null string"); return CATT_E_ARG; }
      cat += cdr3[i0 + i];
      if (cat.size() > 0x7FFF0000ull) { set_error("catt_best_d: CDR3 strings too long"); return CATT_E_LIMIT; }
      off.push_back((int32_t)cat.size());
    }
    int rc;
    if ((rc = c->cl[13].ensure(cat.size() + 16)) || (rc = c->cl[14].ensure(off.size() * sizeof(int32_t))) || (rc = c->cl[15].ensure((size_t)m * sizeof(int32_t)))) return rc;
    CATT_CUDA(cudaMemcpyAsync(c->cl[13].p, cat.data(), cat.size(), cudaMemcpyHostToDevice, c->stream));
    CATT_CUDA(cudaMemcpyAsync(c->cl[14].p, off.data(), off.size() * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    if ((rc = launch_best_d(c->cl[13].as<char>(), c->cl[14].as<int32_t>(), m, c->d_dseq, c->d_doff, (int)c->d_seqs.size(), c->cl[15].as<int32_t>(), c->stream))) return rc;
    CATT_CUDA(cudaMemcpyAsync(out + i0, c->cl[15].p, (size_t)m * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    CATT_CUDA(cudaStreamSynchronize(c->stream));
  }
  return CATT_OK;
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
  FileHits fh;
  fh.start = {0, b->dr.n};
  if ((rc = stage_a_resident(c, b->dr, b->first_ordinal, &fh, &b->info, &launches))) return rc;
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
  if (b->dr.n == 0) return CATT_OK;
  if (v_out) CATT_CUDA(cudaMemcpy(v_out, c->recs_all[0].p, b->dr.n * sizeof(catt_aln), cudaMemcpyDeviceToHost));
  if (j_out) CATT_CUDA(cudaMemcpy(j_out, c->recs_all[1].p, b->dr.n * sizeof(catt_aln), cudaMemcpyDeviceToHost));
  return CATT_OK;
}

void catt_batch_free(catt_batch* b) { if (b) { b->dr.release(); b->quals.release(); b->seqoff.release(); b->names.release(); b->name_off.release(); delete b; } }

int catt_align_shard(catt_ctx* c, int64_t n, const char* seqs, const int64_t* seq_off, int64_t first_ordinal, catt_aln* v_out,
                     catt_aln* j_out) {
  if (!c || n < 0 || (n > 0 && (!seqs || !seq_off))) { set_error("catt_align_shard: bad argument"); return CATT_E_ARG; }
  CATT_CUDA(cudaSetDevice(c->device));
  HostReads hr;
  hr.n = n; hr.seqs = seqs; hr.seq_off = seq_off;
  DeviceReads dr;
  catt_info info{};
  int launches = 0;
  FileHits fh;
  fh.start = {0, n};
  int rc = stage_a_stream(c, hr, first_ordinal, true, false, &dr, nullptr, &fh, &info, &launches);
  if (rc || n == 0) return rc;
  if (v_out) CATT_CUDA(cudaMemcpy(v_out, c->recs_all[0].p, n * sizeof(catt_aln), cudaMemcpyDeviceToHost));
  if (j_out) CATT_CUDA(cudaMemcpy(j_out, c->recs_all[1].p, n * sizeof(catt_aln), cudaMemcpyDeviceToHost));
  return CATT_OK;
}

int catt_assemble(catt_ctx* c, int64_t n, const char* names, const int64_t* name_off, const char* seqs, const int64_t* seq_off,
                  const char* quals, const catt_aln* v_recs, const catt_aln* j_recs, catt_result** out) {
  if (!c || !out || n < 0 || (n > 0 && (!v_recs || !j_recs || !names || !name_off || !seqs || !seq_off))) { set_error("catt_assemble: bad argument"); return CATT_E_ARG; }
  CATT_CUDA(cudaSetDevice(c->device));
  HostReads hr;
  hr.n = n; hr.names = names; hr.name_off = name_off; hr.seqs = seqs; hr.seq_off = seq_off; hr.quals = quals;
  std::unique_ptr<catt_result> R(new catt_result());
  R->info.n_reads = n;
  int64_t vh = 0, jh = 0;
  #pragma omp parallel for reduction(+ : vh, jh) schedule(static)
  for (int64_t i = 0; i < n; i++) { vh += v_recs[i].mapped != 0; jh += j_recs[i].mapped != 0; }
  R->info.v_hits = vh; R->info.j_hits = jh;
  DeviceReads dr;
  int launches = 0;
  FileHits fh;
  fh.start = {0, n};
  DeviceQuals dq;
  int rc = stage_a_stream(c, hr, 0, false, true, &dr, &dq, &fh, &R->info, &launches);      // pack + upload only
  fh.single(n, vh, jh);
  if (rc) return rc;
  if (n) {
    CATT_CUDA(cudaMemcpyAsync(c->recs_all[0].p, v_recs, n * sizeof(catt_aln), cudaMemcpyHostToDevice, c->stream));
    CATT_CUDA(cudaMemcpyAsync(c->recs_all[1].p, j_recs, n * sizeof(catt_aln), cudaMemcpyHostToDevice, c->stream));
    CATT_CUDA(cudaStreamSynchronize(c->stream));
  }
  if (c->keep_records) { R->rec[0].assign(v_recs, v_recs + n); R->rec[1].assign(j_recs, j_recs + n); }
  NamesDev nd;
  if ((rc = upload_names(c, n, names, name_off, &nd, &R->info))) return rc;
  rc = stage_b(c, n, nd, dr.view(), dq, fh, R.get(), &launches);
  if (rc) return rc;
  R->info.kernel_launches = launches;
  *out = R.release();
  return CATT_OK;
}

int catt_shard_counts(const catt_ctx* c, int64_t* n_reads, int64_t* n_words, int32_t* max_len) {
  if (!c || !n_reads || !n_words || !max_len) return CATT_E_ARG;
  *n_reads = c->last_n; *n_words = c->last_words; *max_len = c->last_max_len;
  return CATT_OK;
}

int catt_reads_device(catt_ctx* c, int64_t cap_reads, int64_t cap_words, uint64_t* words_ptr, uint64_t* off_ptr, uint64_t* len_ptr) {
  if (!c || !words_ptr || !off_ptr || !len_ptr || cap_reads < 0 || cap_words < 0) { set_error("catt_reads_device: bad argument"); return CATT_E_ARG; }
  if ((uint64_t)cap_words > 0xFFFFFFF0ull) { set_error("batch too large: packed reads exceed 2^32 words"); return CATT_E_LIMIT; }
  CATT_CUDA(cudaSetDevice(c->device));
  int rc;
  if ((rc = c->rd_words.ensure_keep((size_t)(cap_words + 4) * 4, c->stream)) || (rc = c->rd_off.ensure_keep((size_t)(cap_reads + 2) * 4, c->stream)) ||
      (rc = c->rd_len.ensure_keep((size_t)(cap_reads + 1) * 4, c->stream))) return rc;
  *words_ptr = (uint64_t)(uintptr_t)c->rd_words.p; *off_ptr = (uint64_t)(uintptr_t)c->rd_off.p; *len_ptr = (uint64_t)(uintptr_t)c->rd_len.p;
  return CATT_OK;
}

int catt_records_device(catt_ctx* c, int which, int64_t capacity, uint64_t* device_ptr) {
  if (!c || !device_ptr || which < 0 || which > 1 || capacity < 0) { set_error("catt_records_device: bad argument"); return CATT_E_ARG; }
  CATT_CUDA(cudaSetDevice(c->device));
  int rc = c->recs_all[which].ensure_keep((size_t)std::max<int64_t>(capacity, 1) * sizeof(catt_aln), c->stream);
  if (rc) return rc;
  *device_ptr = (uint64_t)(uintptr_t)c->recs_all[which].p;
  return CATT_OK;
}

namespace {
__global__ void count_mapped_kernel(const catt_aln* __restrict__ recs, int64_t n, unsigned long long* __restrict__ out) {
  unsigned long long mine = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) mine += recs[i].mapped != 0;
  #pragma unroll
  for (int d = 16; d; d >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, d);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(out, mine);
}
}  // namespace

int catt_assemble_device(catt_ctx* c, int64_t n, const char* names, const int64_t* name_off, const char* seqs, const int64_t* seq_off,
                         const char* quals, catt_result** out) {
  if (!c || !out || n < 0 || (n > 0 && (!names || !name_off || !seqs || !seq_off))) { set_error("catt_assemble_device: bad argument"); return CATT_E_ARG; }
  CATT_CUDA(cudaSetDevice(c->device));
  if (c->recs_all[0].cap < (size_t)n * sizeof(catt_aln) || c->recs_all[1].cap < (size_t)n * sizeof(catt_aln)) {
    set_error("catt_assemble_device: the record arrays hold fewer than n records (call catt_records_device first)");
    return CATT_E_ARG;
  }
  HostReads hr;
  hr.n = n; hr.names = names; hr.name_off = name_off; hr.seqs = seqs; hr.seq_off = seq_off; hr.quals = quals;
  std::unique_ptr<catt_result> R(new catt_result());
  R->info.n_reads = n;
  unsigned long long* cnt = c->pin[PIN_CNT].as<unsigned long long>();
  int rc;
  if ((rc = c->sb[SB_CNT].ensure(64 * sizeof(unsigned long long)))) return rc;
  unsigned long long* d_cnt = c->sb[SB_CNT].as<unsigned long long>();
  CATT_CUDA(cudaMemsetAsync(d_cnt, 0, 2 * sizeof(unsigned long long), c->stream));
  if (n) {
    for (int which = 0; which < 2; which++) count_mapped_kernel<<<c->sm_count * 4, 256, 0, c->stream>>>(c->recs_all[which].as<catt_aln>(), n, d_cnt + which);
    CATT_CUDA(cudaGetLastError());
  }
  CATT_CUDA(cudaMemcpyAsync(cnt, d_cnt, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
  CATT_CUDA(cudaStreamSynchronize(c->stream));
  R->info.v_hits = (int64_t)cnt[0]; R->info.j_hits = (int64_t)cnt[1];
  DeviceReads dr;
  DeviceQuals dq;
  int launches = 2;
  FileHits fh;
  fh.start = {0, n};
  // pack + upload only; the records stay where the caller put them (ensure_recs never shrinks or moves a large enough array)
  if ((rc = stage_a_stream(c, hr, 0, false, true, &dr, &dq, &fh, &R->info, &launches))) return rc;
  fh.single(n, R->info.v_hits, R->info.j_hits);
  if (c->keep_records && (rc = download_records(c, n, R.get()))) return rc;
  NamesDev nd;
  if ((rc = upload_names(c, n, names, name_off, &nd, &R->info))) return rc;
  if ((rc = stage_b(c, n, nd, dr.view(), dq, fh, R.get(), &launches))) return rc;
  R->info.kernel_launches = launches;
  *out = R.release();
  return CATT_OK;
}

int catt_assemble_resident(catt_ctx* c, int64_t n, int32_t max_len, const char* names, const int64_t* name_off, const int64_t* seq_off,
                           const char* quals, catt_result** out) {
  if (!c || !out || n < 0 || (n > 0 && (!names || !name_off || !seq_off))) { set_error("catt_assemble_resident: bad argument"); return CATT_E_ARG; }
  CATT_CUDA(cudaSetDevice(c->device));
  const size_t rb = (size_t)n * sizeof(catt_aln);
  if (c->recs_all[0].cap < rb || c->recs_all[1].cap < rb || c->rd_len.cap < (size_t)n * 4 || c->rd_off.cap < (size_t)(n + 1) * 4) {
    set_error("catt_assemble_resident: the context's buffers hold fewer than n reads / records");
    return CATT_E_ARG;
  }
  std::unique_ptr<catt_result> R(new catt_result());
  R->info.n_reads = n;
  int rc;
  if ((rc = c->sb[SB_CNT].ensure(64 * sizeof(unsigned long long)))) return rc;
  unsigned long long* d_cnt = c->sb[SB_CNT].as<unsigned long long>();
  unsigned long long* cnt = c->pin[PIN_CNT].as<unsigned long long>();
  CATT_CUDA(cudaMemsetAsync(d_cnt, 0, 2 * sizeof(unsigned long long), c->stream));
  if (n) {
    for (int which = 0; which < 2; which++) count_mapped_kernel<<<c->sm_count * 4, 256, 0, c->stream>>>(c->recs_all[which].as<catt_aln>(), n, d_cnt + which);
    CATT_CUDA(cudaGetLastError());
  }
  CATT_CUDA(cudaMemcpyAsync(cnt, d_cnt, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
  CATT_CUDA(cudaStreamSynchronize(c->stream));
  R->info.v_hits = (int64_t)cnt[0]; R->info.j_hits = (int64_t)cnt[1];
  DeviceReads dr;
  dr.words = c->rd_words.as<uint32_t>(); dr.off = c->rd_off.as<uint32_t>(); dr.len = c->rd_len.as<int32_t>(); dr.n = n; dr.max_len = max_len;
  DeviceQuals dq;
  if (n > 0) {
    if ((rc = c->rd_seqoff.ensure((size_t)(n + 1) * sizeof(int64_t)))) return rc;
    CATT_CUDA(cudaMemcpyAsync(c->rd_seqoff.p, seq_off, (size_t)(n + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
    dq.seq_off = c->rd_seqoff.as<int64_t>();
    if (c->only_cdr3) { dq.host_quals = quals; dq.host_seq_off = seq_off; }
    else if (quals) {
      if ((rc = c->rd_quals.ensure((size_t)seq_off[n] + 16))) return rc;
      CATT_CUDA(cudaMemcpyAsync(c->rd_quals.p, quals, (size_t)seq_off[n], cudaMemcpyHostToDevice, c->stream));
      dq.quals = c->rd_quals.as<char>();
    }
  }
  int launches = 2;
  FileHits fh;
  fh.single(n, R->info.v_hits, R->info.j_hits);
  if (c->keep_records && (rc = download_records(c, n, R.get()))) return rc;
  NamesDev nd;
  if ((rc = upload_names(c, n, names, name_off, &nd, &R->info))) return rc;
  if ((rc = stage_b(c, n, nd, dr.view(), dq, fh, R.get(), &launches))) return rc;
  R->info.kernel_launches = launches;
  *out = R.release();
  return CATT_OK;
}

// Stage B on a batch that is already resident in HBM and aligned (catt_batch_align): used for the device-resident
// throughput number.  The host buffers are still needed for the QNAMEs and to materialise the Myread strings.
int catt_batch_assemble(catt_ctx* c, catt_batch* b, const char* names, const int64_t* name_off, const char* seqs,
                        const int64_t* seq_off, const char* quals, catt_result** out) {
  if (!c || !b || !out) { set_error("catt_batch_assemble: bad argument"); return CATT_E_ARG; }
  CATT_CUDA(cudaSetDevice(c->device));
  const int64_t n = b->dr.n;
  HostReads hr;
  hr.n = n; hr.names = names; hr.name_off = name_off; hr.seqs = seqs; hr.seq_off = seq_off; hr.quals = quals;
  std::unique_ptr<catt_result> R(new catt_result());
  R->info = b->info;
  int rc = CATT_OK;
  if (c->keep_records) rc = download_records(c, n, R.get());
  if (rc) return rc;
  int launches = (int)b->info.kernel_launches;
  FileHits fh;
  fh.single(n, b->info.v_hits, b->info.j_hits);
  if (!b->have_quals && n > 0) {                       // first assemble of this batch: the quality strings join the resident input
    if ((rc = b->seqoff.ensure((size_t)(n + 1) * sizeof(int64_t)))) return rc;
    CATT_CUDA(cudaMemcpyAsync(b->seqoff.p, seq_off, (size_t)(n + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
    if (quals) {
      if ((rc = b->quals.ensure((size_t)seq_off[n] + 16))) return rc;
      CATT_CUDA(cudaMemcpyAsync(b->quals.p, quals, (size_t)seq_off[n], cudaMemcpyHostToDevice, c->stream));
    }
    if ((rc = b->names.ensure((size_t)name_off[n] + 16)) || (rc = b->name_off.ensure((size_t)(n + 1) * sizeof(int64_t)))) return rc;
    CATT_CUDA(cudaMemcpyAsync(b->names.p, names, (size_t)name_off[n], cudaMemcpyHostToDevice, c->stream));
    CATT_CUDA(cudaMemcpyAsync(b->name_off.p, name_off, (size_t)(n + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
    CATT_CUDA(cudaStreamSynchronize(c->stream));
    b->have_quals = true;
  }
  DeviceQuals dq;
  dq.seq_off = b->seqoff.as<int64_t>(); dq.quals = quals ? b->quals.as<char>() : nullptr;
  NamesDev nd;
  if (n > 0) { nd.s = b->names.as<char>(); nd.beg = b->name_off.as<int64_t>(); nd.end = nd.beg + 1; }
  rc = stage_b(c, n, nd, b->dr.view(), dq, fh, R.get(), &launches);
  if (rc) return rc;
  R->info.kernel_launches = launches;
  *out = R.release();
  return CATT_OK;
}

// Synthetic read generator of SURVEY 8d config 3 (benchmarks and tests; tests/util.py holds the same algorithm in Python):
// read i draws from splitmix64 seeded with seed+i, so any rank can generate its own shard.  Fixed-length output.
int catt_synth_reads(const catt_ctx* c, int64_t n, int64_t first_index, int32_t length, uint64_t seed, double p_err, int32_t with_n,
                     char* seqs, char* quals, char* names, int64_t* name_off) {
  if (!c || !seqs || !quals || length < 20 || length > CATT_MAX_READ_LEN) { set_error("catt_synth_reads: bad argument"); return CATT_E_ARG; }
  const RefSet &V = c->V, &J = c->J;
  const int perr = (int)(p_err * 100000);
  static const char NT[] = "ACGT";
  #pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < n; r++) {
    uint64_t st = seed + (uint64_t)(first_index + r);
    auto below = [&](uint64_t m) { return sm64(st) % m; };
    char tpl[1024];
    int tl = 0;
    const int vi = (int)below(V.names.size()), ji = (int)below(J.names.size());
    const int vlen = V.len[vi] - (int)below(7);
    const int jtrim = (int)below(7);
    const int nins = (int)below(19);
    char ins[32];
    for (int k = 0; k < nins; k++) ins[k] = NT[below(4)];
    for (int k = 0; k < 60; k++) tpl[tl++] = NT[below(4)];
    char post[120];
    for (int k = 0; k < 120; k++) post[k] = NT[below(4)];
    const int v0 = tl;
    for (int k = 0; k < vlen; k++) tpl[tl++] = NT[V.T[V.off[vi] + k]];
    const int v1 = tl;
    for (int k = 0; k < nins; k++) tpl[tl++] = ins[k];
    const int j0 = tl;
    for (int k = jtrim; k < J.len[ji]; k++) tpl[tl++] = NT[J.T[J.off[ji] + k]];
    const int j1 = tl;
    for (int k = 0; k < 120; k++) tpl[tl++] = post[k];
    int start = 0;
    for (int t = 0; t < 64; t++) {
      start = (int)below((uint64_t)std::max(1, tl - length + 1));
      const int en = start + length;
      const int ovv = std::max(0, std::min(en, v1) - std::max(start, v0)), ovj = std::max(0, std::min(en, j1) - std::max(start, j0));
      if (ovv >= 20 || ovj >= 12) break;
    }
    char* s = seqs + r * (int64_t)length;
    char* q = quals + r * (int64_t)length;
    for (int k = 0; k < length; k++) {
      s[k] = start + k < tl ? tpl[start + k] : 'A';
      q[k] = 'I';
      const int x = (int)below(100000);
      if (x < perr) { s[k] = NT[below(4)]; q[k] = '#'; }
      else if (x < 5000 + perr) q[k] = 'G';
    }
    if (with_n && below(50) == 0) s[below((uint64_t)length)] = 'N';
    if (below(2)) {
      auto comp = [](char ch) { return ch == 'A' ? 'T' : ch == 'C' ? 'G' : ch == 'G' ? 'C' : ch == 'T' ? 'A' : 'N'; };
      for (int a = 0, b2 = length - 1; a <= b2; a++, b2--) {
        const char ca = comp(s[a]), cb = comp(s[b2]);
        s[a] = cb; s[b2] = ca;
        const char qa = q[a]; q[a] = q[b2]; q[b2] = qa;
      }
    }
  }
  if (names && name_off) {
    int64_t o = 0;
    for (int64_t r = 0; r < n; r++) { name_off[r] = o; o += snprintf(names + o, 24, "s%lld", (long long)(first_index + r)); }
    name_off[n] = o;
  }
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
  CATT_CUDA(cudaMemsetAsync(ab.counters, 0, N_COUNTERS * sizeof(unsigned long long), c->stream));
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
  rc = launch_finder_raw(dr.view(), c->d_finder, array_version, d_out, c->sm_count, c->stream);
  if (!rc) {
    cudaError_t e = cudaStreamSynchronize(c->stream);
    if (e == cudaSuccess) e = cudaMemcpy(out3, d_out, n * 12, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = cuda_fail(e, "finder D2H", __FILE__, __LINE__);
  }
  cudaFree(d_out); dr.release();
  return rc;
}

int catt_real_score(catt_ctx* c, int which, int64_t n, const char* seqs, const int64_t* seq_off, const int32_t* rid, const int32_t* strand,
                    const int32_t* qb, const int32_t* m_end, const int32_t* rb, int32_t allowance, uint8_t* out) {
  if (!c || !out || !rid || !strand || !qb || !m_end || !rb || which < 0 || which > 1 || n < 0) return CATT_E_ARG;
  if (n == 0) return CATT_OK;
  CATT_CUDA(cudaSetDevice(c->device));
  RefSet& rs = which ? c->J : c->V;
  for (int64_t i = 0; i < n; i++) {
    const int64_t l = seq_off[i + 1] - seq_off[i];
    if (rid[i] < 0 || rid[i] >= (int)rs.names.size() || qb[i] < 0 || m_end[i] < qb[i] + 1 || m_end[i] > l || rb[i] < 0 || rb[i] + (m_end[i] - qb[i] - 1) >= rs.len[rid[i]]) {
      set_error("catt_real_score: record %lld is not a valid alignment record", (long long)i);
      return CATT_E_ARG;
    }
  }
  PackedReads pr;
  int rc = pack_reads(n, seqs, seq_off, &pr);
  if (rc) return rc;
  DeviceReads dr;
  if ((rc = upload_reads(pr, n, &dr, c->stream))) return rc;
  const size_t b = (size_t)n * sizeof(int32_t);
  if ((rc = c->cl[0].ensure(5 * b + (size_t)n + 64))) { dr.release(); return rc; }
  int32_t* d = c->cl[0].as<int32_t>();
  const int32_t* src[5] = {rid, strand, qb, m_end, rb};
  cudaError_t e = cudaSuccess;
  for (int k = 0; k < 5 && e == cudaSuccess; k++) e = cudaMemcpyAsync(d + (size_t)k * n, src[k], b, cudaMemcpyHostToDevice, c->stream);
  uint8_t* d_out = reinterpret_cast<uint8_t*>(d + 5 * (size_t)n);
  if (e == cudaSuccess) rc = launch_real_score(rs.dev, dr.view(), d, d + n, d + 2 * n, d + 3 * n, d + 4 * n, allowance, d_out, c->sm_count, c->stream);
  if (e == cudaSuccess && !rc) e = cudaMemcpyAsync(out, d_out, (size_t)n, cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  dr.release();
  if (e != cudaSuccess) return cuda_fail(e, "catt_real_score", __FILE__, __LINE__);
  return rc;
}

int catt_order_records(catt_ctx* c, int64_t n, const char* names, const int64_t* name_off, const int32_t* read_len, const catt_aln* v_recs,
                       const catt_aln* j_recs, int32_t* items3, int64_t* n_items, int64_t* counts5) {
  if (!c || n < 0 || !n_items || !counts5 || (n > 0 && (!names || !name_off || !read_len || !v_recs || !j_recs || !items3))) { set_error("catt_order_records: bad argument"); return CATT_E_ARG; }
  *n_items = 0;
  for (int k = 0; k < 5; k++) counts5[k] = 0;
  if (n == 0) return CATT_OK;
  CATT_CUDA(cudaSetDevice(c->device));
  int rc;
  if ((rc = ensure_recs(c, n)) || (rc = c->rd_len.ensure((size_t)n * sizeof(int32_t)))) return rc;
  int64_t vh = 0, jh = 0;
  for (int64_t i = 0; i < n; i++) { vh += v_recs[i].mapped != 0; jh += j_recs[i].mapped != 0; }
  CATT_CUDA(cudaMemcpyAsync(c->recs_all[0].p, v_recs, (size_t)n * sizeof(catt_aln), cudaMemcpyHostToDevice, c->stream));
  CATT_CUDA(cudaMemcpyAsync(c->recs_all[1].p, j_recs, (size_t)n * sizeof(catt_aln), cudaMemcpyHostToDevice, c->stream));
  CATT_CUDA(cudaMemcpyAsync(c->rd_len.p, read_len, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
  NamesDev nd;
  if ((rc = upload_names(c, n, names, name_off, &nd, nullptr))) return rc;
  return order_only(c, n, nd, c->rd_len.as<int32_t>(), vh, jh, items3, n_items, counts5);
}

int catt_alu_peak(catt_ctx* c, double* ops, double* ms) {
  if (!c || !ops || !ms) return CATT_E_ARG;
  CATT_CUDA(cudaSetDevice(c->device));
  return alu_peak(c->sm_count, c->stream, ops, ms);
}

}  // extern "C"

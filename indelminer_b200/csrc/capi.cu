// Host side of libindelgpu.so: the C ABI declared in include/indelgpu.h.
// Owns device memory, streams and launches; no torch, no CPU compute path.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/indelgpu.h"
#include "kernels.cuh"
#include "band_dp.cuh"
#include "realign_kernel.cuh"
#include "realign_pipeline.cuh"
#include "task_kernels.cuh"
#include "indel_support.cuh"
#include "indel_support_pack.cuh"

using namespace indelgpu;

// ---------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...)
{
    va_list ap; va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail(INDELGPU_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                                  \
    } while (0)

extern "C" const char* indelgpu_last_error(void) { return g_err; }

extern "C" int indelgpu_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
extern "C" int indelgpu_version(void) { return INDELGPU_VERSION; }

extern "C" void indelgpu_default_params(indelgpu_params* p)
{
    p->klength = 6; p->numgaps = 0; p->maxdelsize = 1000; p->ethreshold = 10;
    p->match = 1; p->mismatch = -10; p->gapopen = 10; p->gapextend = 10;
}

// ---------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------
struct DevBuf {
    void* p = nullptr; size_t cap = 0;
    int ensure(size_t bytes)
    {
        if (bytes <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        if (cudaMalloc(&p, want) != cudaSuccess) { cudaGetLastError(); return fail(INDELGPU_ENOMEM, "cudaMalloc(%zu) failed", want); }
        cap = want;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct WarpPlanFwd { const void* kern; int bytes_per_warp, warps_per_cta, ctas_per_sm; };

struct indelgpu_ctx {
    int device = 0;
    int sms = 0;
    int max_smem_optin = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t st_in = nullptr, st_out = nullptr;   // H2D / D2H streams of the chunked host path
    std::vector<cudaEvent_t> ev_in, ev_k, ev_out;     // per chunk: inputs landed, kernel done, segment count on the host
    void* pinned_counts = nullptr; int pinned_counts_cap = 0;
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;     // around the kernel(s) of the last batch call
    bool timed = false;
    std::vector<WarpPlanFwd> plans;                   // cached launch shapes of the warp-per-read kernels
    indelgpu_params params;
    DevParams P;
    // reference
    DevBuf ref_raw, ref_packed, ref_off, ref_len;
    DevBuf idx_off, idx_pos;                          // resident k-mer index (kmer_index.cuh), built by set_reference for k <= 6
    int64_t idx_blocks = 0;
    int ncontigs = 0;
    int64_t ref_bases = 0;
    // batch staging
    DevBuf in_reads, in_off, in_tid, in_pos, in_rng;
    DevBuf in_seq4, in_boff, in_len, in_flags;        // 4-bit packed batches (indelgpu_realign_batch4), unpacked on the device
    DevBuf out_status, out_nseg, out_rstart, out_segoff, out_segs, out_detail, out_cig1, out_cig2;
    DevBuf chunk_counts;     // chunked host path: segment count after each chunk's kernel, snapshot in stream order
    DevBuf counters;         // bytes: 0 work counter (int) | 8 seg count (u64) | 16 cells (3 x u64) | 40 error flag (int) | 48 algorithmic bytes (u64) | 56 ALIGN cells not swept (u64, band tasks)
    DevBuf scratch;
    DevBuf s_tgt, s_toff, s_qry, s_qoff, s_out, s_ord, s_idx, s_V, s_I, s_F;
    // the two-pass support check (indel_support_pack.cuh): scratch of up to three wavefront launches that share one
    // launch of the walk back
    struct PackSlot { DevBuf dirs, cpl, best; } pk[3];
    int32_t* h_order = nullptr; size_t h_order_cap = 0;            // pinned work list of the support check   // known-indel support check (indel_support.cuh)
    DevBuf p_low, p_aln, p_cig, p_plan, p_flags;      // intermediates of the banded pipeline (realign_pipeline.cuh)
    // task API staging
    DevBuf t_reads, t_roff, t_refs, t_woff, t_packed, t_anchor, t_low, t_up, t_score, t_ends, t_ncig, t_cig, t_script;
    void* pinned_small = nullptr;   // 64 bytes for counter read-back
    int launches = 0;
};

static int* ctr_work(indelgpu_ctx* c) { return c->counters.as<int>(); }
static unsigned long long* ctr_segs(indelgpu_ctx* c) { return reinterpret_cast<unsigned long long*>(c->counters.as<char>() + 8); }
static unsigned long long* ctr_cells(indelgpu_ctx* c) { return reinterpret_cast<unsigned long long*>(c->counters.as<char>() + 16); }
static int* ctr_err(indelgpu_ctx* c) { return reinterpret_cast<int*>(c->counters.as<char>() + 40); }

static int check_params(const indelgpu_params* p)
{
    if (p->klength < 2 || p->klength > 15) return fail(INDELGPU_EINVAL, "klength must be 2..15 (indelminer.c:1028)");
    if (p->numgaps < 0) return fail(INDELGPU_EINVAL, "numgaps must be >= 0");
    if (p->maxdelsize <= 0) return fail(INDELGPU_EINVAL, "maxdelsize must be > 0 (indelminer.c:1027)");
    if (p->ethreshold < 0) return fail(INDELGPU_EINVAL, "ethreshold must be >= 0");
    if (p->gapopen < 0 || p->gapextend < 0) return fail(INDELGPU_EINVAL, "gap penalties must be >= 0");
    return 0;
}

static int ctx_init(indelgpu_ctx* c, int device, const indelgpu_params* p)
{
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (ndev <= 0) return fail(INDELGPU_ECUDA, "no CUDA device visible; libindelgpu has no CPU path");
    if (device < 0 || device >= ndev) return fail(INDELGPU_EINVAL, "device %d out of range (0..%d)", device, ndev - 1);
    CU(cudaSetDevice(device));
    c->device = device;
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    c->sms = prop.multiProcessorCount;
    c->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    if (prop.major < 10) return fail(INDELGPU_ECUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CU(cudaEventCreate(&c->ev_t0));
    CU(cudaEventCreate(&c->ev_t1));
    CU(cudaStreamCreateWithFlags(&c->st_in, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&c->st_out, cudaStreamNonBlocking));
    c->params = *p;
    c->P.k = p->klength; c->P.g = p->numgaps; c->P.maxdel = p->maxdelsize; c->P.ethr = p->ethreshold;
    c->P.match = p->match; c->P.mismatch = p->mismatch; c->P.G = p->gapopen; c->P.H = p->gapextend;
    c->P.kmask = (uint32_t)((1ULL << (2 * p->klength)) - 1ULL);
    if (c->counters.ensure(64)) return INDELGPU_ENOMEM;
    CU(cudaMemsetAsync(c->counters.p, 0, 64, c->stream));
    CU(cudaMallocHost(&c->pinned_small, 64));
    return 0;
}

extern "C" indelgpu_ctx* indelgpu_create(int device, const indelgpu_params* p)
{
    indelgpu_params def;
    if (!p) { indelgpu_default_params(&def); p = &def; }
    if (check_params(p)) return nullptr;
    indelgpu_ctx* c = new indelgpu_ctx();
    if (ctx_init(c, device, p)) { delete c; return nullptr; }
    return c;
}

extern "C" void indelgpu_destroy(indelgpu_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); }
    if (c->st_in) { cudaStreamSynchronize(c->st_in); cudaStreamDestroy(c->st_in); }
    if (c->st_out) { cudaStreamSynchronize(c->st_out); cudaStreamDestroy(c->st_out); }
    if (c->ev_t0) cudaEventDestroy(c->ev_t0);
    if (c->ev_t1) cudaEventDestroy(c->ev_t1);
    for (cudaEvent_t e : c->ev_in) cudaEventDestroy(e);
    for (cudaEvent_t e : c->ev_k) cudaEventDestroy(e);
    for (cudaEvent_t e : c->ev_out) cudaEventDestroy(e);
    if (c->pinned_counts) cudaFreeHost(c->pinned_counts);
    if (c->h_order) cudaFreeHost(c->h_order);
    DevBuf* all[] = {&c->ref_raw, &c->ref_packed, &c->ref_off, &c->ref_len, &c->idx_off, &c->idx_pos, &c->in_reads, &c->in_off, &c->in_tid, &c->in_seq4, &c->in_boff, &c->in_len, &c->in_flags,
                     &c->in_pos, &c->in_rng, &c->out_status, &c->out_nseg, &c->out_rstart, &c->out_segoff,
                     &c->out_segs, &c->out_detail, &c->out_cig1, &c->out_cig2, &c->counters, &c->chunk_counts, &c->scratch,
                     &c->p_low, &c->p_aln, &c->p_cig, &c->p_plan, &c->p_flags,
                     &c->s_tgt, &c->s_toff, &c->s_qry, &c->s_qoff, &c->s_out, &c->s_ord, &c->s_idx, &c->s_V, &c->s_I, &c->s_F, &c->pk[0].dirs, &c->pk[0].cpl, &c->pk[0].best, &c->pk[1].dirs, &c->pk[1].cpl, &c->pk[1].best, &c->pk[2].dirs, &c->pk[2].cpl, &c->pk[2].best,
                     &c->t_reads, &c->t_roff, &c->t_refs, &c->t_woff, &c->t_packed, &c->t_anchor, &c->t_low,
                     &c->t_up, &c->t_score, &c->t_ends, &c->t_ncig, &c->t_cig, &c->t_script};
    for (DevBuf* b : all) b->release();
    if (c->pinned_small) cudaFreeHost(c->pinned_small);
    delete c;
}

extern "C" int indelgpu_device(const indelgpu_ctx* c) { return c ? c->device : -1; }
extern "C" int indelgpu_sm_count(const indelgpu_ctx* c) { return c ? c->sms : 0; }
extern "C" int indelgpu_last_launch_count(const indelgpu_ctx* c) { return c ? c->launches : 0; }

extern "C" void* indelgpu_host_alloc(size_t bytes)
{
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); fail(INDELGPU_ENOMEM, "cudaMallocHost(%zu) failed", bytes); return nullptr; }
    return p;
}
extern "C" void indelgpu_host_free(void* p) { if (p) cudaFreeHost(p); }

// ---------------------------------------------------------------------------------------
// reference upload
// ---------------------------------------------------------------------------------------
static int pack_device(indelgpu_ctx* c, const uint8_t* d_raw, uint32_t* d_packed, int64_t nwords)
{
    if (nwords <= 0) return 0;
    int blocks = (int)std::min<int64_t>((nwords + 255) / 256, (int64_t)c->sms * 16);
    pack_reference_kernel<<<blocks, 256, 0, c->stream>>>(d_raw, d_packed, nwords);
    c->launches++;
    CU(cudaGetLastError());
    return 0;
}

extern "C" int indelgpu_set_reference(indelgpu_ctx* c, int32_t ncontigs, const char* const* sequences,
                                      const int64_t* lengths)
{
    if (!c || ncontigs <= 0 || !sequences || !lengths) return fail(INDELGPU_EINVAL, "set_reference: bad arguments");
    CU(cudaSetDevice(c->device));
    std::vector<int64_t> off(ncontigs), len(ncontigs);
    int64_t total = 0;
    for (int i = 0; i < ncontigs; i++) {
        if (lengths[i] < 0 || lengths[i] > 0x7FFFFFFF) return fail(INDELGPU_ELIMIT, "contig %d length %lld not in [0, 2^31)", i, (long long)lengths[i]);
        off[i] = total; len[i] = lengths[i];
        total += (lengths[i] + 63) / 64 * 64 + 64;       // 64-base aligned start + slack
    }
    total += 64;
    if (c->ref_raw.ensure((size_t)total)) return INDELGPU_ENOMEM;
    if (c->ref_packed.ensure((size_t)(total / 16 + 4) * 4)) return INDELGPU_ENOMEM;
    if (c->ref_off.ensure(sizeof(int64_t) * ncontigs) || c->ref_len.ensure(sizeof(int64_t) * ncontigs)) return INDELGPU_ENOMEM;
    CU(cudaMemsetAsync(c->ref_raw.p, 0, (size_t)total, c->stream));
    for (int i = 0; i < ncontigs; i++)
        if (len[i] > 0) CU(cudaMemcpyAsync(c->ref_raw.as<uint8_t>() + off[i], sequences[i], (size_t)len[i], cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->ref_off.p, off.data(), sizeof(int64_t) * ncontigs, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->ref_len.p, len.data(), sizeof(int64_t) * ncontigs, cudaMemcpyHostToDevice, c->stream));
    c->launches = 0;
    if (pack_device(c, c->ref_raw.as<uint8_t>(), c->ref_packed.as<uint32_t>(), total / 16)) return INDELGPU_ECUDA;
    CU(cudaMemsetAsync(c->ref_packed.as<uint32_t>() + total / 16, 0, 16, c->stream));
    // the k-mer index of the whole reference (kmer_index.cuh): 8 * 4^k + 2 * 4096 bytes per 4096 bases.
    // EXPERIMENTAL, opt-in with INDELGPU_INDEX=1: measured on B200 (profiles/r02_index_vote.md) the index vote is exact
    // but executes MORE instructions per read than the window scan (6.2 k vs 4.7 k warp instructions; 8.2 ms vs 6.7 ms
    // per 1 Mi reads) -- per-bucket bookkeeping costs more than the scan's ~5 instructions per window position saves.
    c->idx_blocks = 0;
    if (c->P.k <= kIdxMaxK && getenv("INDELGPU_INDEX") != nullptr && atoi(getenv("INDELGPU_INDEX")) > 0) {
        const int64_t nblocks = (total + kIdxBlock - 1) / kIdxBlock;
        const size_t off_bytes = (size_t)nblocks * 8 * ((size_t)1 << (2 * c->P.k)), pos_bytes = (size_t)nblocks * 2 * kIdxBlock;
        size_t free_b = 0, total_b = 0;
        CU(cudaMemGetInfo(&free_b, &total_b));
        const size_t need = (off_bytes > c->idx_off.cap ? off_bytes : 0) + (pos_bytes > c->idx_pos.cap ? pos_bytes : 0);
        if (need + (need >> 2) + ((size_t)2 << 30) < free_b) {                 // leave room for the batches
            if (c->ref_packed.ensure((size_t)(nblocks * (kIdxBlock / 16) + 4) * 4)) return INDELGPU_ENOMEM;   // no-op: total / 16 + 4 words cover it
            if (c->idx_off.ensure(off_bytes) || c->idx_pos.ensure(pos_bytes)) return INDELGPU_ENOMEM;
            const int blocks = (int)std::min<int64_t>(nblocks, (int64_t)c->sms * 8);
            build_kmer_index_kernel<<<blocks, 256, 0, c->stream>>>(c->ref_packed.as<uint32_t>(), nblocks, c->P.k, c->P.kmask,
                                                                   c->idx_off.as<uint2>(), c->idx_pos.as<uint16_t>());
            c->launches++;
            CU(cudaGetLastError());
            c->idx_blocks = nblocks;
        }
    }
    CU(cudaStreamSynchronize(c->stream));
    c->ncontigs = ncontigs; c->ref_bases = total;
    return 0;
}

// ---------------------------------------------------------------------------------------
// batched attempt_pe_alignment
// ---------------------------------------------------------------------------------------
extern "C" int64_t indelgpu_seg_bound(int32_t n, int64_t total_read_bases)
{
    // <= (M+2) ops per CIGAR, two CIGARs + one I + one D per read
    return 2 * total_read_bases + 8LL * n + 16;
}


// The dynamic shared-memory limit is a property of the FUNCTION on a device, shared by every context and
// host thread of the process: it is only ever raised to the device's opt-in maximum, never set to what
// one launch needs, so that concurrent contexts with different window sizes cannot lower it under
// each other's launches.
template <class Kern>
static cudaError_t allow_max_smem(indelgpu_ctx* c, Kern kern)
{
    cudaFuncAttributes fa;
    const cudaError_t e = cudaFuncGetAttributes(&fa, kern);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, c->max_smem_optin - (int)fa.sharedSizeBytes);
}

// Persistent warp-per-read kernels: pick the CTA size that puts the most warps on an SM given the
// per-warp shared-memory slice.
typedef WarpPlanFwd WarpPlan;

template <class Kern>
static int plan_warps(indelgpu_ctx* c, Kern kern, int bytes_per_warp, int* warps_per_cta, int* ctas_per_sm)
{
    // the answer only depends on (kernel, slice size): remember it, the runtime queries are not free
    for (const WarpPlan& w : c->plans)
        if (w.kern == (const void*)kern && w.bytes_per_warp == bytes_per_warp) {
            *warps_per_cta = w.warps_per_cta; *ctas_per_sm = w.ctas_per_sm;
            return 0;
        }
    static const int cand[] = {8, 7, 6, 5, 4, 3, 2, 1};       // __launch_bounds__(256)
    CU(allow_max_smem(c, kern));
    int best = 0;
    *warps_per_cta = 0; *ctas_per_sm = 0;
    for (int w : cand) {
        const long long smem = (long long)w * bytes_per_warp;
        if (smem > c->max_smem_optin) continue;
        int occ = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, w * 32, (size_t)smem));
        if (occ * w > best) { best = occ * w; *warps_per_cta = w; *ctas_per_sm = occ; }
    }
    if (best == 0) return fail(INDELGPU_ELIMIT, "window/read sizes need %d bytes of shared memory per warp (limit %d): range1 + maxdelsize or the read length is too large",
                               bytes_per_warp, c->max_smem_optin);
    if (const char* e = getenv("INDELGPU_MAX_WARPS_PER_SM")) {   // occupancy experiments only
        const int cap = atoi(e);
        if (cap > 0) {
            *warps_per_cta = std::min(*warps_per_cta, cap);
            *ctas_per_sm = std::max(1, std::min(*ctas_per_sm, cap / *warps_per_cta));
        }
    }
    c->plans.push_back(WarpPlan{(const void*)kern, bytes_per_warp, *warps_per_cta, *ctas_per_sm});
    return 0;
}

// a batch in the BAM's 4-bit form, on the device (fused -g 0 kernel only; banded runs unpack it first)
struct Packed4Dev { const uint8_t* seq4; const int64_t* byte_off; const uint8_t* flags; };

static void fill_realign_args(indelgpu_ctx* c, RealignArgs& a, const indelgpu_batch* d_in, indelgpu_result* d_out,
                              unsigned long long* d_seg_count, int max_read, int max_numdiag, const WarpLayout& L,
                              const int32_t* d_read_len = nullptr, const Packed4Dev* p4 = nullptr)
{
    a.read_len = d_read_len;
    a.seq4 = p4 ? p4->seq4 : nullptr; a.byte_off = p4 ? p4->byte_off : nullptr; a.rflags = p4 ? p4->flags : nullptr;
    a.P = c->P;
    a.ref.raw = c->ref_raw.as<uint8_t>(); a.ref.packed = c->ref_packed.as<uint32_t>();
    a.ref.contig_off = c->ref_off.as<int64_t>(); a.ref.contig_len = c->ref_len.as<int64_t>(); a.ref.ncontigs = c->ncontigs;
    a.n = d_in->n; a.reads = d_in->read_bases; a.read_off = d_in->read_off;
    a.tid = d_in->tid; a.position = d_in->position; a.range1 = d_in->range1;
    a.status = d_out->status; a.nseg = d_out->nseg; a.rstart = d_out->rstart; a.seg_off = d_out->seg_off;
    a.segs = d_out->segs; a.seg_capacity = d_out->seg_capacity; a.seg_count = d_seg_count;
    a.detail = d_out->detail; a.cigar1 = d_out->cigar1; a.cigar2 = d_out->cigar2; a.cigar_stride = d_out->cigar_stride;
    a.work_counter = ctr_work(c); a.cell_totals = ctr_cells(c); a.error_flag = ctr_err(c);
    a.max_read = max_read; a.max_numdiag = max_numdiag; a.L = L;
    a.idx.rec = c->idx_off.as<uint2>(); a.idx.pos = c->idx_pos.as<uint16_t>(); a.idx.nblocks = c->idx_blocks; a.idx.k = c->idx_blocks ? c->P.k : 0;
    a.scratch.base = nullptr; a.scratch.stride = 0; a.scratch.max_band = 0; a.scratch.max_rows = 0;
}

// -g N > 0: vote / DP / vote / DP / combine (realign_pipeline.cuh)
static int launch_pipeline(indelgpu_ctx* c, const indelgpu_batch* d_in, int max_read, int max_numdiag, const WarpLayout& L,
                           indelgpu_result* d_out, unsigned long long* d_seg_count, cudaStream_t st, bool keep_totals,
                           const int32_t* d_read_len)
{
    const int n = d_in->n;
    const int max_band = c->P.g + 1;
    PipeBufs p;
    p.cig_stride = 2 * max_read + max_band + 4;
    if (c->p_low.ensure(8 * (size_t)n) || c->p_aln.ensure(2 * sizeof(Aln) * (size_t)n) || c->p_plan.ensure(sizeof(Plan) * (size_t)n) ||
        c->p_flags.ensure(4 * (size_t)n) || c->p_cig.ensure(8 * (size_t)n * (size_t)p.cig_stride)) return INDELGPU_ENOMEM;
    p.low = c->p_low.as<int32_t>(); p.aln = c->p_aln.as<Aln>(); p.cig = c->p_cig.as<uint32_t>();
    p.plan = c->p_plan.as<Plan>(); p.flags = c->p_flags.as<int32_t>();

    RealignArgs a;
    fill_realign_args(c, a, d_in, d_out, d_seg_count, max_read, max_numdiag, L, d_read_len);

    void (*vk)(RealignArgs, PipeBufs, int);
    if (L.hist_bits == 8) vk = L.direct ? pipe_vote_kernel<true, 8> : pipe_vote_kernel<false, 8>;
    else                  vk = L.direct ? pipe_vote_kernel<true, 16> : pipe_vote_kernel<false, 16>;
    int wpc = 0, occ = 0;
    if (int rc = plan_warps(c, vk, L.total, &wpc, &occ)) return rc;
    const int vblocks = (int)std::min<long long>((long long)c->sms * occ, std::max(1, (n + wpc - 1) / wpc));

    const int mb = 2 * max_band;
    // (no shared-memory carve-out preference: asking for the largest L1 made the driver pick a carve-out too
    //  small for four resident CTAs and cost 30 %; the default leaves the L1-resident work arrays enough room)
    int docc = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&docc, pipe_dp_kernel, 128, 0));
    if (docc < 1) return fail(INDELGPU_ELIMIT, "banded DP kernel does not fit on an SM");
    const int dblocks = (int)std::min<long long>((long long)c->sms * docc, (n + 127) / 128);
    const long long ints = band_scratch_ints(mb, max_read);
    if (c->scratch.ensure((size_t)ints * 4 * 128 * (size_t)dblocks)) return INDELGPU_ENOMEM;
    a.scratch.base = c->scratch.as<int>(); a.scratch.stride = ints; a.scratch.max_band = mb; a.scratch.max_rows = max_read;

    const int cper = (4 * p.cig_stride + 8) * 4 + 256;
    int cw = 8;
    while (cw > 1 && (size_t)cw * cper > (size_t)c->max_smem_optin) cw >>= 1;
    if ((size_t)cw * cper > (size_t)c->max_smem_optin) return fail(INDELGPU_ELIMIT, "reads too long for the combine kernel's shared memory");
    CU(allow_max_smem(c, pipe_combine_kernel));
    int cocc = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cocc, pipe_combine_kernel, cw * 32, (size_t)cw * cper));
    if (cocc < 1) return fail(INDELGPU_ELIMIT, "combine kernel does not fit on an SM");
    const int cblocks = (int)std::min<long long>((long long)c->sms * cocc, std::max(1, (n + cw - 1) / cw));

    if (keep_totals) CU(cudaMemsetAsync(c->counters.p, 0, 4, st));
    else {
        CU(cudaMemsetAsync(c->counters.p, 0, 64, st));
        if (d_seg_count != ctr_segs(c)) CU(cudaMemsetAsync(d_seg_count, 0, 8, st));
    }
    for (int round = 0; round < 2; round++) {
        vk<<<vblocks, wpc * 32, (size_t)wpc * L.total, st>>>(a, p, round);
        CU(cudaGetLastError());
        pipe_dp_kernel<<<dblocks, 128, 0, st>>>(a, p, round);
        CU(cudaGetLastError());
    }
    pipe_combine_kernel<<<cblocks, cw * 32, (size_t)cw * cper, st>>>(a, p);
    CU(cudaGetLastError());
    c->launches += 5;
    return 0;
}

// keep_totals: a later chunk of the same batch -- only the work counter is reset, the segment
// counter, the cell totals and the error flag keep accumulating
static int launch_realign(indelgpu_ctx* c, const indelgpu_batch* d_in, int max_read, int max_range1,
                          indelgpu_result* d_out, unsigned long long* d_seg_count, cudaStream_t st,
                          bool keep_totals = false, const int32_t* d_read_len = nullptr, const Packed4Dev* p4 = nullptr)
{
    if (c->ncontigs <= 0) return fail(INDELGPU_EINVAL, "realign: no reference uploaded (indelgpu_set_reference)");
    if (max_read <= 0 || max_read > 65000) return fail(INDELGPU_ELIMIT, "read length %d outside 1..65000", max_read);
    const long long nd = 2LL * ((long long)max_range1 + c->P.maxdel) + max_read + 2;
    if (nd > (1 << 20)) return fail(INDELGPU_ELIMIT, "window of %lld diagonals exceeds the kernel limit", nd);
    const int max_numdiag = (int)nd;
    static const bool force_split = getenv("INDELGPU_SPLIT") != nullptr;     // experiment: -g 0 through the kernel pipeline
    const bool banded = c->P.g > 0 || (force_split && !p4);
    // the banded pipeline only votes with this layout (its CIGARs live in HBM), so it takes the compact form too
    const long long vd = std::max(2LL * max_range1, (long long)max_range1 + c->P.maxdel) + max_read + 2;
    if (p4 && (c->P.k > 6 || c->idx_blocks > 0)) return fail(INDELGPU_EINVAL, "internal: only the direct-table window-scan kernel expands 4-bit reads");
    const bool indexed = !banded && c->idx_blocks > 0;            // -g 0, k <= 6: the fused kernel votes through the resident index
    if (p4 && banded) return fail(INDELGPU_EINVAL, "internal: 4-bit batches reach the banded pipeline unpacked");
    const WarpLayout L = make_warp_layout(c->P, max_read, max_numdiag, (int)vd, 0, indexed ? 1 : 0, p4 ? 1 : 0);
    if (p4 && ((uintptr_t)p4->seq4 & 15) != 0) return fail(INDELGPU_EINVAL, "seq4 must be 16-byte aligned on the device (TMA bulk copies)");
    if (((uintptr_t)d_in->read_bases & 15) != 0) return fail(INDELGPU_EINVAL, "read_bases must be 16-byte aligned on the device (TMA bulk copies)");
    if (banded) return launch_pipeline(c, d_in, max_read, max_numdiag, L, d_out, d_seg_count, st, keep_totals, d_read_len);
    void (*kern)(RealignArgs);
    if (p4)               kern = L.hist_bits == 8 ? realign_kernel<true, 8, false, true> : realign_kernel<true, 16, false, true>;
    else if (L.indexed)   kern = L.hist_bits == 8 ? realign_kernel<true, 8, true> : realign_kernel<true, 16, true>;
    else if (L.hist_bits == 8) kern = L.direct ? realign_kernel<true, 8, false> : realign_kernel<false, 8, false>;
    else                  kern = L.direct ? realign_kernel<true, 16, false> : realign_kernel<false, 16, false>;
    int wpc = 0, occ = 0;
    if (int rc = plan_warps(c, kern, L.total, &wpc, &occ)) return rc;
    const int blocks = (int)std::min<long long>((long long)c->sms * occ, std::max(1, (d_in->n + wpc - 1) / wpc));

    RealignArgs a;
    fill_realign_args(c, a, d_in, d_out, d_seg_count, max_read, max_numdiag, L, d_read_len, p4);

    if (keep_totals) CU(cudaMemsetAsync(c->counters.p, 0, 4, st));
    else {
        CU(cudaMemsetAsync(c->counters.p, 0, 64, st));
        if (d_seg_count != ctr_segs(c)) CU(cudaMemsetAsync(d_seg_count, 0, 8, st));
    }
    kern<<<blocks, wpc * 32, (size_t)wpc * L.total, st>>>(a);
    c->launches++;
    CU(cudaGetLastError());
    return 0;
}

extern "C" int indelgpu_realign_batch_device(indelgpu_ctx* c, const indelgpu_batch* d_in, int32_t max_read_len,
                                             int32_t max_range1, indelgpu_result* d_out, int64_t* d_seg_count,
                                             void* stream)
{
    if (!c || !d_in || !d_out || !d_seg_count) return fail(INDELGPU_EINVAL, "realign_batch_device: NULL argument");
    if (d_in->n < 0) return fail(INDELGPU_EINVAL, "negative batch size");
    CU(cudaSetDevice(c->device));
    c->launches = 0;
    if (d_in->n == 0) return 0;
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    return launch_realign(c, d_in, max_read_len, max_range1, d_out, reinterpret_cast<unsigned long long*>(d_seg_count), st);
}

extern "C" double indelgpu_last_kernel_ms(indelgpu_ctx* c)
{
    if (!c || !c->timed) return -1.0;
    if (cudaSetDevice(c->device) != cudaSuccess || cudaEventSynchronize(c->ev_t1) != cudaSuccess) return -1.0;
    float ms = -1.0f;
    if (cudaEventElapsedTime(&ms, c->ev_t0, c->ev_t1) != cudaSuccess) return -1.0;
    return (double)ms;
}

// INT32 issue-rate micro-benchmark (the denominator of the banded-DP roofline is
// measured on the box, it is not in MEASURED_PEAKS.json): dependency-free add + max chains.
__global__ void int32_peak_kernel(int iters, int seed, int* out)
{
    int v[8];
#pragma unroll
    for (int u = 0; u < 8; u++) v[u] = (int)threadIdx.x * (u + 1) + seed;
    const int b = (int)blockIdx.x - seed;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) v[u] = max(v[u] + b, i - u);        // one add and one max per statement
    }
    int r = 0;
#pragma unroll
    for (int u = 0; u < 8; u++) r ^= v[u];
    if (r == 0x7FFFFFFF) out[0] = r;
}

extern "C" int indelgpu_int32_peak(indelgpu_ctx* c, double* gops)
{
    if (!c || !gops) return fail(INDELGPU_EINVAL, "int32_peak: NULL argument");
    CU(cudaSetDevice(c->device));
    if (c->t_ncig.ensure(16)) return INDELGPU_ENOMEM;
    const int iters = 1 << 14, blocks = c->sms * 16, threads = 256;
    double best = 0;
    for (int rep = 0; rep < 5; rep++) {
        CU(cudaEventRecord(c->ev_t0, c->stream));
        int32_peak_kernel<<<blocks, threads, 0, c->stream>>>(iters, rep, c->t_ncig.as<int>());
        CU(cudaEventRecord(c->ev_t1, c->stream));
        CU(cudaGetLastError());
        CU(cudaEventSynchronize(c->ev_t1));
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, c->ev_t0, c->ev_t1));
        const double ops = 2.0 * 8.0 * (double)iters * (double)blocks * (double)threads;   // add + max
        if (rep > 0) best = std::max(best, ops / (ms * 1e-3) / 1e9);
    }
    *gops = best;
    return 0;
}

extern "C" int indelgpu_last_counters(indelgpu_ctx* c, int64_t out[4])
{
    if (!c || !out) return fail(INDELGPU_EINVAL, "last_counters: NULL argument");
    CU(cudaSetDevice(c->device));
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(c->pinned_small, c->counters.p, 64, cudaMemcpyDeviceToHost));
    memcpy(out, (char*)c->pinned_small + 16, 24);
    memcpy(out + 3, (char*)c->pinned_small + 48, 8);
    return 0;
}

// ALIGN cells of the last indelgpu_band_align_batch that were counted (SURVEY.md 8d defines Gc by what the reference
// sweeps) but NOT swept here, because the unique-diagonal shortcut proved the all-REP script optimal (band_dp.cuh).
// executed cells = forward + reverse + align - this.
extern "C" int indelgpu_last_shortcut_cells(indelgpu_ctx* c, int64_t* out)
{
    if (!c || !out) return fail(INDELGPU_EINVAL, "last_shortcut_cells: NULL argument");
    CU(cudaSetDevice(c->device));
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(c->pinned_small, c->counters.p, 64, cudaMemcpyDeviceToHost));
    memcpy(out, (char*)c->pinned_small + 56, 8);
    return 0;
}

// Error flag of the last batch launched on this context, after waiting for it: 0 none, 1 at least one read
// was rejected (status 7), 2 the segment buffer overflowed (seg_off / nseg are set, the words are not),
// 3 a TMA bulk copy never completed.  indelgpu_realign_batch checks it itself; callers of the
// device-pointer entry point must ask.
extern "C" int indelgpu_last_error_flag(indelgpu_ctx* c)
{
    if (!c) return fail(INDELGPU_EINVAL, "last_error_flag: NULL context");
    CU(cudaSetDevice(c->device));
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(c->pinned_small, c->counters.p, 64, cudaMemcpyDeviceToHost));
    int err; memcpy(&err, (char*)c->pinned_small + 40, 4);
    return err;
}

// BAM 4-bit bases -> ASCII on the device (bit2char, readaln.c:4-17: 1 A, 2 C, 4 G, 8 T, 15 N; anything else is an
// input the reference stops on: error flag 1), reverse-complemented when the read's flag asks for it
// (reverse_complement_string, sequences.c:204-220, on A C G T N).  One warp per read; read i lands at 2 * byte_off[i].
__global__ void unpack_reads4_kernel(int n, const uint8_t* __restrict__ seq4, const int64_t* __restrict__ byte_off,
                                     const int32_t* __restrict__ len, const uint8_t* __restrict__ flags,
                                     uint8_t* __restrict__ out, int64_t* __restrict__ out_off, int* error_flag)
{
    const int lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    #pragma unroll 1
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += nwarps) {
        const int64_t b0 = byte_off[i];
        const int L = len[i];
        const bool rc = (flags[i] & 1u) != 0u;
        uint8_t* dst = out + 2 * b0;
        if (lane == 0) { out_off[i] = 2 * b0; if (i == n - 1) out_off[n] = 2 * byte_off[n]; }
        #pragma unroll 1
        for (int j = lane; 2 * j < L; j += 32) {
            const uint32_t byte = seq4[b0 + j];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int t = 2 * j + h;
                if (t >= L) break;
                const uint32_t code = h == 0 ? (byte >> 4) : (byte & 15u);      // bam1_seqi: high nibble first
                uint8_t fwd, rev;
                switch (code) {
                    case 1: fwd = 'A'; rev = 'T'; break;
                    case 2: fwd = 'C'; rev = 'G'; break;
                    case 4: fwd = 'G'; rev = 'C'; break;
                    case 8: fwd = 'T'; rev = 'A'; break;
                    case 15: fwd = 'N'; rev = 'N'; break;
                    default: fwd = 'N'; rev = 'N'; atomicExch(error_flag, 1); break;
                }
                if (rc) dst[L - 1 - t] = rev; else dst[t] = fwd;
            }
        }
    }
}

// host view of a batch in either input form
struct HostBatch {
    int n;
    const uint8_t* bases; const int64_t* off;                    // ASCII: read i = bases[off[i] .. off[i+1])
    const int32_t* len; const uint8_t* flags; bool packed4;      // 4-bit: bases = BAM nibbles, off = byte offsets
    const int32_t* tid; const int32_t* position; const int32_t* range1;
};

static int realign_batch_impl(indelgpu_ctx* c, const HostBatch* h, indelgpu_result* o);

extern "C" int indelgpu_realign_batch(indelgpu_ctx* c, const indelgpu_batch* b, indelgpu_result* o)
{
    if (!c || !b || !o) return fail(INDELGPU_EINVAL, "realign_batch: NULL argument");
    if (b->n > 0 && (!b->read_bases || !b->read_off || !b->tid || !b->position || !b->range1)) return fail(INDELGPU_EINVAL, "realign_batch: NULL buffer");
    const HostBatch h = {b->n, b->read_bases, b->read_off, nullptr, nullptr, false, b->tid, b->position, b->range1};
    return realign_batch_impl(c, &h, o);
}

extern "C" int indelgpu_realign_batch4(indelgpu_ctx* c, const indelgpu_batch4* b, indelgpu_result* o)
{
    if (!c || !b || !o) return fail(INDELGPU_EINVAL, "realign_batch4: NULL argument");
    if (b->n > 0 && (!b->seq4 || !b->byte_off || !b->len || !b->flags || !b->tid || !b->position || !b->range1)) return fail(INDELGPU_EINVAL, "realign_batch4: NULL buffer");
    if (o->detail || o->cigar1 || o->cigar2) return fail(INDELGPU_EINVAL, "realign_batch4: debug outputs are only available through indelgpu_realign_batch");
    const HostBatch h = {b->n, b->seq4, b->byte_off, b->len, b->flags, true, b->tid, b->position, b->range1};
    return realign_batch_impl(c, &h, o);
}

static int realign_batch_impl(indelgpu_ctx* c, const HostBatch* h, indelgpu_result* o)
{
    const int n = h->n;
    if (n < 0) return fail(INDELGPU_EINVAL, "negative batch size");
    o->seg_count = 0;
    c->launches = 0;
    if (n == 0) return 0;
    if (!o->status || !o->nseg || !o->rstart || !o->seg_off || !o->segs)
        return fail(INDELGPU_EINVAL, "realign_batch: NULL buffer");
    const bool p4 = h->packed4;
    CU(cudaSetDevice(c->device));
    // sizes are validated (and the kernel limits derived) chunk by chunk, right before a chunk is enqueued,
    // so that the first copies do not wait for a pass over the whole batch
    auto scan_range = [&](int lo, int hi, int* mr, int* mg) -> int {
        int max_read = 0, max_range = 0;
        for (int i = lo; i < hi; i++) {
            const int64_t len = p4 ? (int64_t)h->len[i] : h->off[i + 1] - h->off[i];
            if (len <= 0 || len > 65000) return fail(INDELGPU_ELIMIT, "read %d has length %lld (must be 1..65000)", i, (long long)len);
            if (p4 && h->off[i + 1] - h->off[i] != (len + 1) / 2) return fail(INDELGPU_EINVAL, "read %d: %lld bases do not fill %lld bytes", i, (long long)len, (long long)(h->off[i + 1] - h->off[i]));
            max_read = std::max(max_read, (int)len);
            if (h->range1[i] < 0) return fail(INDELGPU_EINVAL, "read %d: negative range", i);
            max_range = std::max(max_range, h->range1[i]);
        }
        *mr = max_read; *mg = max_range;
        return 0;
    };
    // device ASCII bytes: a 4-bit read of L bases occupies 2 * ceil(L / 2) of them
    const int64_t nbytes_in = h->off[n] - h->off[0];
    const int64_t nbases = p4 ? 2 * nbytes_in : nbytes_in;
    if (h->off[0] != 0) return fail(INDELGPU_EINVAL, "the first read offset must be 0");
    const int64_t segcap = std::min<int64_t>(o->seg_capacity, indelgpu_seg_bound(n, nbases));
    cudaStream_t st = c->stream;
    if (c->in_reads.ensure((size_t)nbases + 16) || c->in_off.ensure(8 * (size_t)(n + 1)) || c->in_tid.ensure(4 * (size_t)n) ||
        c->in_pos.ensure(4 * (size_t)n) || c->in_rng.ensure(4 * (size_t)n) || c->out_status.ensure(4 * (size_t)n) ||
        c->out_nseg.ensure(4 * (size_t)n) || c->out_rstart.ensure(4 * (size_t)n) || c->out_segoff.ensure(8 * (size_t)n) ||
        c->out_segs.ensure(4 * (size_t)std::max<int64_t>(segcap, 1)))
        return INDELGPU_ENOMEM;
    if (p4 && (c->in_seq4.ensure((size_t)nbytes_in + 16) || c->in_boff.ensure(8 * (size_t)(n + 1)) || c->in_len.ensure(4 * (size_t)n) ||
               c->in_flags.ensure((size_t)n + 16))) return INDELGPU_ENOMEM;
    if (o->detail && c->out_detail.ensure(sizeof(indelgpu_detail) * (size_t)n)) return INDELGPU_ENOMEM;
    const size_t cigbytes = 4 * (size_t)n * (size_t)std::max(o->cigar_stride, 0);
    if (o->cigar1 && c->out_cig1.ensure(cigbytes + 4)) return INDELGPU_ENOMEM;
    if (o->cigar2 && c->out_cig2.ensure(cigbytes + 4)) return INDELGPU_ENOMEM;

    indelgpu_batch din;
    din.n = n;
    din.read_bases = c->in_reads.as<uint8_t>(); din.read_off = c->in_off.as<int64_t>();
    din.tid = c->in_tid.as<int32_t>(); din.position = c->in_pos.as<int32_t>(); din.range1 = c->in_rng.as<int32_t>();
    indelgpu_result dout = *o;
    dout.status = c->out_status.as<int32_t>(); dout.nseg = c->out_nseg.as<int32_t>(); dout.rstart = c->out_rstart.as<int32_t>();
    dout.seg_off = c->out_segoff.as<int64_t>(); dout.segs = c->out_segs.as<uint32_t>(); dout.seg_capacity = segcap;
    dout.detail = o->detail ? c->out_detail.as<indelgpu_detail>() : nullptr;
    dout.cigar1 = o->cigar1 ? c->out_cig1.as<uint32_t>() : nullptr;
    dout.cigar2 = o->cigar2 ? c->out_cig2.as<uint32_t>() : nullptr;

    // host -> device copies of reads [c0, c1) on stream `si`; 4-bit batches are unpacked on `sk` once `ev` says they landed
    auto enqueue_inputs = [&](int c0, int c1, cudaStream_t si) -> int {
        const int m = c1 - c0;
        const int64_t b0 = h->off[c0], b1 = h->off[c1];
        if (!p4) {
            CU(cudaMemcpyAsync(c->in_off.as<int64_t>() + c0, h->off + c0, 8 * (size_t)(m + 1), cudaMemcpyHostToDevice, si));
            CU(cudaMemcpyAsync(c->in_reads.as<uint8_t>() + b0, h->bases + b0, (size_t)(b1 - b0), cudaMemcpyHostToDevice, si));
        } else {
            CU(cudaMemcpyAsync(c->in_boff.as<int64_t>() + c0, h->off + c0, 8 * (size_t)(m + 1), cudaMemcpyHostToDevice, si));
            CU(cudaMemcpyAsync(c->in_seq4.as<uint8_t>() + b0, h->bases + b0, (size_t)(b1 - b0), cudaMemcpyHostToDevice, si));
            CU(cudaMemcpyAsync(c->in_len.as<int32_t>() + c0, h->len + c0, 4 * (size_t)m, cudaMemcpyHostToDevice, si));
            CU(cudaMemcpyAsync(c->in_flags.as<uint8_t>() + c0, h->flags + c0, (size_t)m, cudaMemcpyHostToDevice, si));
        }
        CU(cudaMemcpyAsync(c->in_tid.as<int32_t>() + c0, h->tid + c0, 4 * (size_t)m, cudaMemcpyHostToDevice, si));
        CU(cudaMemcpyAsync(c->in_pos.as<int32_t>() + c0, h->position + c0, 4 * (size_t)m, cudaMemcpyHostToDevice, si));
        CU(cudaMemcpyAsync(c->in_rng.as<int32_t>() + c0, h->range1 + c0, 4 * (size_t)m, cudaMemcpyHostToDevice, si));
        return 0;
    };
    // -g 0: the fused kernel expands the 4-bit reads itself; the banded pipeline reads ASCII, so the batch is unpacked first
    const bool fused4 = p4 && c->P.g == 0 && c->P.k <= 6 && c->idx_blocks == 0;
    auto enqueue_unpack = [&](int c0, int c1, cudaStream_t sk) -> int {
        if (!p4 || fused4) return 0;
        const int m = c1 - c0;
        const int blocks = (int)std::min<long long>((long long)c->sms * 8, (m + 7) / 8);
        unpack_reads4_kernel<<<blocks, 256, 0, sk>>>(m, c->in_seq4.as<uint8_t>(), c->in_boff.as<int64_t>() + c0, c->in_len.as<int32_t>() + c0,
                                                      c->in_flags.as<uint8_t>() + c0, c->in_reads.as<uint8_t>(), c->in_off.as<int64_t>() + c0,
                                                      ctr_err(c));
        c->launches++;
        CU(cudaGetLastError());
        return 0;
    };
    const int32_t* d_len = p4 ? c->in_len.as<int32_t>() : nullptr;
    Packed4Dev dev4 = {c->in_seq4.as<uint8_t>(), c->in_boff.as<int64_t>(), c->in_flags.as<uint8_t>()};

    // Large batches without debug outputs are cut into chunks so that the H2D copy of chunk i+1 and
    // the D2H copy of chunk i-1 overlap the kernel of chunk i (three streams, two events per chunk).
    // Chunks share the reference, the segment allocator and the counters; read offsets stay absolute.
    int64_t segs_copied = 0;                         // segment words the chunked path has already brought back
    const bool debug_out = o->detail || o->cigar1 || o->cigar2;
    int kChunkReads = 1 << 18;
    if (const char* e = getenv("INDELGPU_CHUNK_READS")) { const int v = atoi(e); if (v >= 64) kChunkReads = v; }
    // chunk boundaries: the first chunks are small (1/8, 1/4, 1/2 of a chunk) so that the first kernel starts
    // after a short copy instead of a full chunk's; then full chunks
    std::vector<int> cuts(1, 0);
    if (!(debug_out || n < kChunkReads)) {
        // chunk boundaries: the first chunks are small (1/8, 1/4, 1/2 of a chunk) so that the first kernel starts after a
        // short copy instead of a full chunk's; the last ones shrink the same way (1/2, 1/4, 1/8) so that little is left
        // to copy back after the last kernel; equal chunks of at most kChunkReads in between
        const int k8 = std::max(64, kChunkReads / 8), k4 = std::max(64, kChunkReads / 4), k2 = std::max(64, kChunkReads / 2);
        const bool down = n >= 2 * (k8 + k4 + k2) + kChunkReads / 2;
        const int tail = down ? k8 + k4 + k2 : 0;
        for (int step : {k8, k4, k2}) if (cuts.back() < n - tail) cuts.push_back(std::min(n - tail, cuts.back() + step));
        const int mid = n - tail - cuts.back();
        if (mid > 0) {
            const int pieces = (mid + kChunkReads - 1) / kChunkReads;
            const int a0 = cuts.back();
            for (int q = 1; q <= pieces; q++) cuts.push_back(a0 + (int)((long long)mid * q / pieces));
        }
        if (down) { cuts.push_back(n - k4 - k8); cuts.push_back(n - k8); cuts.push_back(n); }
    } else cuts.push_back(n);
    const int nchunks = (int)cuts.size() - 1;
    if (nchunks > 1) {
        while ((int)c->ev_in.size() < nchunks) {
            cudaEvent_t e1, e2;
            CU(cudaEventCreateWithFlags(&e1, cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&e2, cudaEventDisableTiming));
            c->ev_in.push_back(e1); c->ev_k.push_back(e2);
        }
        if (!c->pinned_counts || c->pinned_counts_cap < nchunks) {
            if (c->pinned_counts) cudaFreeHost(c->pinned_counts);
            c->pinned_counts = nullptr; c->pinned_counts_cap = 0;
            CU(cudaMallocHost(&c->pinned_counts, 8 * (size_t)(nchunks + 8)));
            c->pinned_counts_cap = nchunks + 8;
        }
        while ((int)c->ev_out.size() < nchunks) {
            cudaEvent_t e;
            CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            c->ev_out.push_back(e);
        }
        unsigned long long* counts = reinterpret_cast<unsigned long long*>(c->pinned_counts);
        const bool verbose = getenv("INDELGPU_VERBOSE") != nullptr;
        const auto t_begin = std::chrono::steady_clock::now();
        if (verbose) CU(cudaEventRecord(c->ev_t0, st));
        if (c->chunk_counts.ensure(8 * (size_t)(nchunks + 8))) return INDELGPU_ENOMEM;
        unsigned long long* d_counts = c->chunk_counts.as<unsigned long long>();
        for (int ch = 0; ch < nchunks; ch++) {
            const int c0 = cuts[ch], c1 = cuts[ch + 1], m = c1 - c0;
            int max_read = 0, max_range = 0;
            if (int rcs = scan_range(c0, c1, &max_read, &max_range)) { cudaDeviceSynchronize(); return rcs; }
            if (int rci = enqueue_inputs(c0, c1, c->st_in)) { cudaDeviceSynchronize(); return rci; }
            CU(cudaEventRecord(c->ev_in[ch], c->st_in));
            CU(cudaStreamWaitEvent(st, c->ev_in[ch], 0));
            // (the unpack kernel's error flag write must not be wiped by the first chunk's counter reset: reset first)
            if (ch == 0) CU(cudaMemsetAsync(c->counters.p, 0, 64, st));
            if (int rcu = enqueue_unpack(c0, c1, st)) { cudaDeviceSynchronize(); return rcu; }
            indelgpu_batch dc = din; dc.n = m;
            dc.read_off = din.read_off + c0; dc.tid = din.tid + c0; dc.position = din.position + c0; dc.range1 = din.range1 + c0;
            indelgpu_result rc2 = dout;
            rc2.status = dout.status + c0; rc2.nseg = dout.nseg + c0; rc2.rstart = dout.rstart + c0; rc2.seg_off = dout.seg_off + c0;
            Packed4Dev dc4 = {dev4.seq4, dev4.byte_off + c0, dev4.flags + c0};
            int rcl = launch_realign(c, &dc, max_read, max_range, &rc2, ctr_segs(c), st, true, d_len ? d_len + c0 : nullptr, fused4 ? &dc4 : nullptr);
            if (rcl) { cudaDeviceSynchronize(); return rcl; }
            // The segment allocator keeps counting while chunk ch+1 runs, and the words of a read are written
            // one read later than they are allocated (realign_kernel's pend_* flush).  The count that bounds
            // chunk ch's words is therefore snapshot ON THE KERNEL STREAM, between kernel ch and kernel ch+1:
            // every word below it has been written when ev_k[ch] fires.
            CU(cudaMemcpyAsync(d_counts + ch, ctr_segs(c), 8, cudaMemcpyDeviceToDevice, st));
            CU(cudaEventRecord(c->ev_k[ch], st));
            CU(cudaStreamWaitEvent(c->st_out, c->ev_k[ch], 0));
            CU(cudaMemcpyAsync(counts + ch, d_counts + ch, 8, cudaMemcpyDeviceToHost, c->st_out));
            CU(cudaEventRecord(c->ev_out[ch], c->st_out));
            CU(cudaMemcpyAsync(o->status + c0, dout.status + c0, 4 * (size_t)m, cudaMemcpyDeviceToHost, c->st_out));
            CU(cudaMemcpyAsync(o->nseg + c0, dout.nseg + c0, 4 * (size_t)m, cudaMemcpyDeviceToHost, c->st_out));
            CU(cudaMemcpyAsync(o->rstart + c0, dout.rstart + c0, 4 * (size_t)m, cudaMemcpyDeviceToHost, c->st_out));
            CU(cudaMemcpyAsync(o->seg_off + c0, dout.seg_off + c0, 8 * (size_t)m, cudaMemcpyDeviceToHost, c->st_out));
        }
        const auto t_enq = std::chrono::steady_clock::now();
        if (verbose) CU(cudaEventRecord(c->ev_t1, st));
        // segment words: the allocator is monotonic and the chunks' kernels run in order, so the words of
        // chunk ch are [count after ch-1, count after ch); copy each range as soon as its count is known,
        // while the later chunks are still being realigned
        unsigned long long prev = 0;
        for (int ch = 0; ch < nchunks; ch++) {
            CU(cudaEventSynchronize(c->ev_out[ch]));
            const unsigned long long cur = std::min<unsigned long long>(counts[ch], (unsigned long long)segcap);
            if (cur > prev) CU(cudaMemcpyAsync(o->segs + prev, dout.segs + prev, 4 * (size_t)(cur - prev), cudaMemcpyDeviceToHost, c->st_out));
            prev = std::max(prev, cur);
        }
        segs_copied = (int64_t)prev;
        CU(cudaStreamSynchronize(c->st_out));
        if (verbose) {
            const auto t_end = std::chrono::steady_clock::now();
            float kms = 0;
            cudaEventSynchronize(c->ev_t1);
            cudaEventElapsedTime(&kms, c->ev_t0, c->ev_t1);
            fprintf(stderr, "libindelgpu: realign_batch: %d chunks; host enqueued everything after %.2f ms; kernel stream busy from its first to its last "
                            "command %.2f ms; results on the host after %.2f ms\n", nchunks,
                    std::chrono::duration<double, std::milli>(t_enq - t_begin).count(), (double)kms,
                    std::chrono::duration<double, std::milli>(t_end - t_begin).count());
        }
    } else {
        int max_read = 0, max_range = 0;
        if (int rcs = scan_range(0, n, &max_read, &max_range)) return rcs;
        if (int rci = enqueue_inputs(0, n, st)) return rci;
        CU(cudaMemsetAsync(c->counters.p, 0, 64, st));
        if (int rcu = enqueue_unpack(0, n, st)) return rcu;
        if (o->cigar1) CU(cudaMemsetAsync(c->out_cig1.p, 0, cigbytes, st));
        if (o->cigar2) CU(cudaMemsetAsync(c->out_cig2.p, 0, cigbytes, st));
        int rc = launch_realign(c, &din, max_read, max_range, &dout, ctr_segs(c), st, true, d_len, fused4 ? &dev4 : nullptr);
        if (rc) return rc;
        CU(cudaMemcpyAsync(o->status, dout.status, 4 * (size_t)n, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(o->nseg, dout.nseg, 4 * (size_t)n, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(o->rstart, dout.rstart, 4 * (size_t)n, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(o->seg_off, dout.seg_off, 8 * (size_t)n, cudaMemcpyDeviceToHost, st));
        if (o->detail) CU(cudaMemcpyAsync(o->detail, dout.detail, sizeof(indelgpu_detail) * (size_t)n, cudaMemcpyDeviceToHost, st));
        if (o->cigar1) CU(cudaMemcpyAsync(o->cigar1, dout.cigar1, cigbytes, cudaMemcpyDeviceToHost, st));
        if (o->cigar2) CU(cudaMemcpyAsync(o->cigar2, dout.cigar2, cigbytes, cudaMemcpyDeviceToHost, st));
    }
    CU(cudaMemcpyAsync(c->pinned_small, c->counters.p, 64, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    unsigned long long segcount; int err;
    memcpy(&segcount, (char*)c->pinned_small + 8, 8);
    memcpy(&err, (char*)c->pinned_small + 40, 4);
    if (err == 3) return fail(INDELGPU_ECUDA, "a TMA bulk copy never completed (mbarrier wait timed out)");
    if (err == 2 || (int64_t)segcount > segcap) return fail(INDELGPU_ELIMIT, "segment buffer too small: need %llu words, have %lld", segcount, (long long)segcap);
    o->seg_count = (int64_t)segcount;
    if ((int64_t)segcount > segs_copied)
        CU(cudaMemcpyAsync(o->segs + segs_copied, dout.segs + segs_copied, 4 * (size_t)((int64_t)segcount - segs_copied), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (err == 1) return fail(INDELGPU_ELIMIT, "at least one read was rejected (status %d): bad contig/position, window assert of alignment.c:548-553 or a size limit", ST_ASSERT);
    return 0;
}

// ---------------------------------------------------------------------------------------
// task API
// ---------------------------------------------------------------------------------------
static int upload_tasks(indelgpu_ctx* c, int n, const uint8_t* h_reads, const int64_t* h_read_off,
                        const uint8_t* h_refs, const int64_t* h_ref_off, int* max_read, int* max_win)
{
    if (h_read_off[0] != 0 || h_ref_off[0] != 0) return fail(INDELGPU_EINVAL, "offset arrays must start at 0");
    int mr = 0, mw = 0;
    for (int i = 0; i < n; i++) {
        const int64_t m = h_read_off[i + 1] - h_read_off[i], w = h_ref_off[i + 1] - h_ref_off[i];
        if (m <= 0 || m > 65000) return fail(INDELGPU_ELIMIT, "task %d: read length %lld outside 1..65000", i, (long long)m);
        if (w <= 0 || w > (1 << 20)) return fail(INDELGPU_ELIMIT, "task %d: window length %lld outside 1..2^20", i, (long long)w);
        mr = std::max(mr, (int)m); mw = std::max(mw, (int)w);
    }
    *max_read = mr; *max_win = mw;
    const int64_t nb = h_read_off[n], nw = h_ref_off[n];
    if (c->t_reads.ensure((size_t)nb + 16) || c->t_roff.ensure(8 * (size_t)(n + 1)) ||
        c->t_refs.ensure((size_t)nw + 64) || c->t_woff.ensure(8 * (size_t)(n + 1))) return INDELGPU_ENOMEM;
    cudaStream_t st = c->stream;
    CU(cudaMemcpyAsync(c->t_reads.p, h_reads, (size_t)nb, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(c->t_roff.p, h_read_off, 8 * (size_t)(n + 1), cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(c->t_refs.as<uint8_t>() + nw, 0, 64, st));
    CU(cudaMemcpyAsync(c->t_refs.p, h_refs, (size_t)nw, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(c->t_woff.p, h_ref_off, 8 * (size_t)(n + 1), cudaMemcpyHostToDevice, st));
    return 0;
}

extern "C" int indelgpu_find_best_band_batch(indelgpu_ctx* c, int32_t n, const uint8_t* h_reads,
                                             const int64_t* h_read_off, const uint8_t* h_refs,
                                             const int64_t* h_ref_off, const int32_t* h_anchor_rel,
                                             int32_t* h_low, int32_t* h_up)
{
    if (!c || n < 0 || !h_reads || !h_read_off || !h_refs || !h_ref_off || !h_anchor_rel || !h_low || !h_up)
        return fail(INDELGPU_EINVAL, "find_best_band_batch: bad argument");
    c->launches = 0;
    if (n == 0) return 0;
    CU(cudaSetDevice(c->device));
    int max_read, max_win;
    int rc = upload_tasks(c, n, h_reads, h_read_off, h_refs, h_ref_off, &max_read, &max_win);
    if (rc) return rc;
    cudaStream_t st = c->stream;
    const int64_t nw = h_ref_off[n];
    const int64_t words = (nw + 63) / 16 + 1;
    if (c->t_packed.ensure(4 * (size_t)(words + 4)) || c->t_anchor.ensure(4 * (size_t)n) ||
        c->t_low.ensure(4 * (size_t)n) || c->t_up.ensure(4 * (size_t)n)) return INDELGPU_ENOMEM;
    CU(cudaMemcpyAsync(c->t_anchor.p, h_anchor_rel, 4 * (size_t)n, cudaMemcpyHostToDevice, st));
    if (pack_device(c, c->t_refs.as<uint8_t>(), c->t_packed.as<uint32_t>(), words)) return INDELGPU_ECUDA;

    const int max_numdiag = max_win + max_read + 4;
    const WarpLayout L = make_warp_layout(c->P, max_read, max_numdiag, max_numdiag, 0);
    void (*vkern)(TaskArgs);
    if (L.hist_bits == 8) vkern = L.direct ? vote_tasks_kernel<true, 8> : vote_tasks_kernel<false, 8>;
    else                  vkern = L.direct ? vote_tasks_kernel<true, 16> : vote_tasks_kernel<false, 16>;
    int wpc = 0, occ = 0;
    if (int rc2 = plan_warps(c, vkern, L.total, &wpc, &occ)) return rc2;
    TaskArgs a; memset(&a, 0, sizeof(a));
    a.P = c->P; a.n = n;
    a.reads = c->t_reads.as<uint8_t>(); a.read_off = c->t_roff.as<int64_t>();
    a.refs = c->t_refs.as<uint8_t>(); a.ref_off = c->t_woff.as<int64_t>();
    a.packed = c->t_packed.as<uint32_t>(); a.anchor_rel = c->t_anchor.as<int32_t>();
    a.low = c->t_low.as<int32_t>(); a.up = c->t_up.as<int32_t>();
    a.work_counter = ctr_work(c); a.cell_totals = ctr_cells(c); a.error_flag = ctr_err(c);
    a.max_read = max_read; a.max_numdiag = max_numdiag; a.L = L;
    CU(cudaMemsetAsync(c->counters.p, 0, 64, st));
    const int blocks = (int)std::min<long long>((long long)c->sms * occ, (n + wpc - 1) / wpc);
    vkern<<<blocks, wpc * 32, (size_t)wpc * L.total, st>>>(a);
    c->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(h_low, a.low, 4 * (size_t)n, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h_up, a.up, 4 * (size_t)n, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(c->pinned_small, c->counters.p, 64, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    int err; memcpy(&err, (char*)c->pinned_small + 40, 4);
    if (err == 3) return fail(INDELGPU_ECUDA, "a TMA bulk copy never completed (mbarrier wait timed out)");
    if (err) return fail(INDELGPU_ELIMIT, "find_best_band_batch: a task violates numdiagonals > numgaps (alignment.c:405) or a size limit");
    return 0;
}

extern "C" int indelgpu_band_align_batch(indelgpu_ctx* c, int32_t n, const uint8_t* h_reads,
                                         const int64_t* h_read_off, const uint8_t* h_refs,
                                         const int64_t* h_ref_off, const int32_t* h_low, const int32_t* h_up,
                                         int32_t* h_score, int32_t* h_ends, int32_t* h_ncigar,
                                         uint32_t* h_cigar, int32_t cigar_stride, int32_t* h_script,
                                         int32_t script_stride, int64_t* h_cells)
{
    if (!c || n < 0 || !h_reads || !h_read_off || !h_refs || !h_ref_off || !h_low || !h_up || !h_score || !h_ends || !h_ncigar)
        return fail(INDELGPU_EINVAL, "band_align_batch: bad argument");
    c->launches = 0;
    if (h_cells) h_cells[0] = h_cells[1] = h_cells[2] = 0;
    if (n == 0) return 0;
    CU(cudaSetDevice(c->device));
    int max_read, max_win;
    int rc = upload_tasks(c, n, h_reads, h_read_off, h_refs, h_ref_off, &max_read, &max_win);
    if (rc) return rc;
    cudaStream_t st = c->stream;
    int max_band = 1;
    for (int i = 0; i < n; i++) {
        const int M = (int)(h_read_off[i + 1] - h_read_off[i]), N = (int)(h_ref_off[i + 1] - h_ref_off[i]);
        const int lo = std::max(-M, h_low[i]), hi = std::min(N, h_up[i]);
        if (hi - lo + 1 < 1) return fail(INDELGPU_EINVAL, "task %d: low > up is unacceptable! low:%d up:%d (localalign.c:74-77)", i, lo, hi);
        max_band = std::max(max_band, hi - lo + 1);
    }
    if (cigar_stride < 0) cigar_stride = 0;
    if (script_stride < 0) script_stride = 0;
    if (c->t_low.ensure(4 * (size_t)n) || c->t_up.ensure(4 * (size_t)n) || c->t_score.ensure(4 * (size_t)n) ||
        c->t_ends.ensure(16 * (size_t)n) || c->t_ncig.ensure(4 * (size_t)n) ||
        c->t_cig.ensure(4 * (size_t)n * (size_t)cigar_stride + 4) ||
        c->t_script.ensure(4 * (size_t)n * (size_t)script_stride + 4)) return INDELGPU_ENOMEM;
    CU(cudaMemcpyAsync(c->t_low.p, h_low, 4 * (size_t)n, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(c->t_up.p, h_up, 4 * (size_t)n, cudaMemcpyHostToDevice, st));

    TaskArgs a; memset(&a, 0, sizeof(a));
    a.P = c->P; a.n = n;
    a.reads = c->t_reads.as<uint8_t>(); a.read_off = c->t_roff.as<int64_t>();
    a.refs = c->t_refs.as<uint8_t>(); a.ref_off = c->t_woff.as<int64_t>();
    a.low = c->t_low.as<int32_t>(); a.up = c->t_up.as<int32_t>();
    a.score = c->t_score.as<int32_t>(); a.ends = c->t_ends.as<int32_t>(); a.ncigar = c->t_ncig.as<int32_t>();
    a.cigar = h_cigar ? c->t_cig.as<uint32_t>() : nullptr; a.cigar_stride = cigar_stride;
    a.script = h_script ? c->t_script.as<int32_t>() : nullptr; a.script_stride = script_stride;
    a.work_counter = ctr_work(c); a.cell_totals = ctr_cells(c); a.error_flag = ctr_err(c);
    a.max_read = max_read; a.max_numdiag = 0;
    a.scratch.base = nullptr; a.scratch.stride = 0; a.scratch.max_band = 0; a.scratch.max_rows = 0;
    CU(cudaMemsetAsync(c->counters.p, 0, 64, st));
    if (h_cigar) CU(cudaMemsetAsync(c->t_cig.p, 0, 4 * (size_t)n * (size_t)cigar_stride, st));
    c->timed = false;
    if (max_band > 1) {
        // one alignment per thread on lane-interleaved scratch (band_tasks_kernel)
        const int need = 2 * max_read + max_band + 4;
        if (h_cigar && cigar_stride < need)
            return fail(INDELGPU_EINVAL, "band_align_batch: cigar_stride %d too small for band %d (need %d)", cigar_stride, max_band, need);
        if (!h_cigar) {                                  // the kernel always needs somewhere to build the CIGAR
            if (c->t_cig.ensure(4 * (size_t)n * (size_t)need + 4)) return INDELGPU_ENOMEM;
            a.cigar = c->t_cig.as<uint32_t>(); a.cigar_stride = need;
        }
        const int mb = 2 * max_band;
        int occ = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, band_tasks_kernel, 128, 0));
        if (occ < 1) return fail(INDELGPU_ELIMIT, "band kernel does not fit on an SM");
        const int blocks = (int)std::min<long long>((long long)c->sms * occ, (n + 127) / 128);
        const long long ints = band_scratch_ints(mb, max_read);
        if (c->scratch.ensure((size_t)ints * 4 * 128 * (size_t)blocks)) return INDELGPU_ENOMEM;
        a.scratch.base = c->scratch.as<int>(); a.scratch.stride = ints; a.scratch.max_band = mb; a.scratch.max_rows = max_read;
        CU(cudaEventRecord(c->ev_t0, st));
        band_tasks_kernel<<<blocks, 128, 0, st>>>(a);
        CU(cudaEventRecord(c->ev_t1, st));
        c->timed = true;
    } else if (getenv("INDELGPU_ALIGN1_WARP") == nullptr) {
        // bands of one diagonal, one alignment per thread (the warp-per-task kernel stays behind INDELGPU_ALIGN1_WARP=1)
        int occ = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, align1_tasks_kernel, 128, 0));
        if (occ < 1) return fail(INDELGPU_ELIMIT, "align kernel does not fit on an SM");
        const int blocks = (int)std::min<long long>((long long)c->sms * occ, (n + 127) / 128);
        CU(cudaEventRecord(c->ev_t0, st));
        align1_tasks_kernel<<<blocks, 128, 0, st>>>(a);
        CU(cudaEventRecord(c->ev_t1, st));
        c->timed = true;
    } else {
        const SmemLayout L = make_layout(max_read, 0);
        if (L.total > c->max_smem_optin - 1024) return fail(INDELGPU_ELIMIT, "read too long for shared memory (%d bytes)", L.total);
        CU(allow_max_smem(c, align_tasks_kernel));
        int occ = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, align_tasks_kernel, 32, L.total));
        if (occ < 1) return fail(INDELGPU_ELIMIT, "align kernel does not fit on an SM");
        const int blocks = (int)std::min<long long>((long long)c->sms * occ, n);
        CU(cudaEventRecord(c->ev_t0, st));
        align_tasks_kernel<<<blocks, 32, L.total, st>>>(a);
        CU(cudaEventRecord(c->ev_t1, st));
        c->timed = true;
    }
    c->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(h_score, a.score, 4 * (size_t)n, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h_ends, a.ends, 16 * (size_t)n, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h_ncigar, a.ncigar, 4 * (size_t)n, cudaMemcpyDeviceToHost, st));
    if (h_cigar) CU(cudaMemcpyAsync(h_cigar, a.cigar, 4 * (size_t)n * (size_t)cigar_stride, cudaMemcpyDeviceToHost, st));
    if (h_script) CU(cudaMemcpyAsync(h_script, a.script, 4 * (size_t)n * (size_t)script_stride, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(c->pinned_small, c->counters.p, 64, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    int err; memcpy(&err, (char*)c->pinned_small + 40, 4);
    if (h_cells) memcpy(h_cells, (char*)c->pinned_small + 16, 24);
    if (err) return fail(INDELGPU_ELIMIT, "band_align_batch: a task exceeds a size limit");
    return 0;
}

// ---------------------------------------------------------------------------------------
// known-indel support check (row f1)
// ---------------------------------------------------------------------------------------
extern "C" int indelgpu_indel_support_batch(indelgpu_ctx* c, int32_t n, const uint8_t* h_targets,
                                            const int64_t* h_target_off, const uint8_t* h_queries,
                                            const int64_t* h_query_off, int32_t* h_subs, int32_t* h_indels,
                                            int32_t* h_aligned, int64_t* h_cells)
{
    if (!c || n < 0 || !h_targets || !h_target_off || !h_queries || !h_query_off || !h_subs || !h_indels || !h_aligned)
        return fail(INDELGPU_EINVAL, "indel_support_batch: bad argument");
    c->launches = 0;
    if (h_cells) *h_cells = 0;
    if (n == 0) return 0;
    CU(cudaSetDevice(c->device));
    if (h_target_off[0] != 0 || h_query_off[0] != 0) return fail(INDELGPU_EINVAL, "offset arrays must start at 0");
    // the sequence copies start first: the host pass below (validation, classes, counting sort) runs under them
    const int64_t nt = h_target_off[n], nq = h_query_off[n];
    if (nt < 0 || nq < 0) return fail(INDELGPU_EINVAL, "indel_support_batch: negative offsets");
    cudaStream_t st = c->stream;
    if (c->s_tgt.ensure((size_t)nt + 16) || c->s_toff.ensure(8 * (size_t)(n + 1)) || c->s_qry.ensure((size_t)nq + 16) ||
        c->s_qoff.ensure(8 * (size_t)(n + 1)) || c->s_out.ensure(12 * (size_t)n) || c->s_ord.ensure(4 * (size_t)n)) return INDELGPU_ENOMEM;
    if (nt > 0) CU(cudaMemcpyAsync(c->s_tgt.p, h_targets, (size_t)nt, cudaMemcpyHostToDevice, st));
    if (nq > 0) CU(cudaMemcpyAsync(c->s_qry.p, h_queries, (size_t)nq, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(c->s_toff.p, h_target_off, 8 * (size_t)(n + 1), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(c->s_qoff.p, h_query_off, 8 * (size_t)(n + 1), cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(c->counters.p, 0, 64, st));
    // Six work lists for the two wavefront kernels, each sorted by target length (counting sort): three classes by
    // target length (8, 16 or 32 lanes per pair) x query short enough for the 16-bit two-pass kernel
    // (indel_support_pack.cuh: every short read) or not (indel_support.cuh); the rest, one pair per thread.
    // Small batches (annotate mode scores the ~30 reads of one variant per call) stay on round 1's single-launch wavefront
    // kernel: the two-pass kernels need a few thousand pairs to fill the GPU and win from ~6 000 pairs on (measured:
    // 30 pairs 0.20 against 0.28 ms per call, 8 000 pairs 0.63 against 0.60 ms, 65 536 pairs 4.1 against 3.0 ms).
    // INDELGPU_SUPPORT_PACK_MIN overrides the threshold (0: always two-pass); INDELGPU_SUPPORT_NOPACK=1: never.
    const char* pm = getenv("INDELGPU_SUPPORT_PACK_MIN");
    const int pack_min_pairs = pm ? atoi(pm) : 4096;
    const bool no_pack = getenv("INDELGPU_SUPPORT_NOPACK") != nullptr || n < pack_min_pairs;
    const int pack_max_query = no_pack ? -1 : (int)kPackMaxQuery;
    std::vector<int32_t> slow;
    int max1 = 0, max2 = 0;
    long long cells = 0;
    const int NB = kWaveMaxTarget + 1;                             // bucket = (wave ? NB : 0) + target length
    std::vector<int32_t> bucket(2 * NB + 1, 0);
    int pack_max2[3] = {0, 0, 0};                                  // longest query per packed class: sizes its shared memory
    for (int i = 0; i < n; i++) {
        const int64_t l1 = h_target_off[i + 1] - h_target_off[i], l2 = h_query_off[i + 1] - h_query_off[i];
        if (l1 < 0 || l2 < 0 || l1 > 8000 || l2 > 8000) {
            cudaStreamSynchronize(st);                         // the copies read the caller's buffers
            return fail(INDELGPU_ELIMIT, "task %d: lengths %lld x %lld outside 0..8000", i, (long long)l1, (long long)l2);
        }
        cells += l1 * l2;
        if (l1 > kWaveMaxTarget || l2 > kWaveMaxQuery) { slow.push_back(i); max1 = std::max(max1, (int)l1); max2 = std::max(max2, (int)l2); }
        else if (l2 <= pack_max_query) { bucket[l1 + 1]++; int& m = pack_max2[l1 <= 128 ? 0 : l1 <= 256 ? 1 : 2]; m = std::max(m, (int)l2); }
        else bucket[NB + l1 + 1]++;
    }
    const int nslow = (int)slow.size(), nfast = n - nslow;
    for (int l = 0; l < 2 * NB; l++) bucket[l + 1] += bucket[l];
    // list boundaries: packed <= 128 | <= 256 | <= 512 | wave <= 128 | <= 256 | <= 512
    const int list_end[6] = {bucket[128 + 1], bucket[256 + 1], bucket[NB], bucket[NB + 128 + 1], bucket[NB + 256 + 1], bucket[2 * NB]};
    if (c->h_order_cap < (size_t)std::max(nfast, 1)) {           // pinned: the copy below must not wait for a staging pass
        if (c->h_order) cudaFreeHost(c->h_order);
        c->h_order = nullptr; c->h_order_cap = 0;
        CU(cudaMallocHost((void**)&c->h_order, 4 * (size_t)std::max(n, 1)));
        c->h_order_cap = (size_t)std::max(n, 1);
    }
    int32_t* order = c->h_order;
    {
        std::vector<int32_t> at(bucket.begin(), bucket.end() - 1);
        for (int i = 0; i < n; i++) {
            const int64_t l1 = h_target_off[i + 1] - h_target_off[i], l2 = h_query_off[i + 1] - h_query_off[i];
            if (l1 <= kWaveMaxTarget && l2 <= kWaveMaxQuery) order[at[(l2 <= pack_max_query ? 0 : NB) + l1]++] = i;
        }
    }
    if (nfast > 0) CU(cudaMemcpyAsync(c->s_ord.p, order, 4 * (size_t)nfast, cudaMemcpyHostToDevice, st));
    int32_t* d_subs = c->s_out.as<int32_t>();
    int32_t* d_indels = d_subs + n;
    int32_t* d_aligned = d_indels + n;
    CU(cudaEventRecord(c->ev_t0, st));
    // the 16-bit two-pass kernels: a wavefront pass that leaves two direction bits per cell in a scratch buffer, then
    // the walk back with one thread per pair.  A class is cut into launches whose scratch stays under ~2 GB.
    const int pack_shape = getenv("INDELGPU_PACK_SHAPE") ? atoi(getenv("INDELGPU_PACK_SHAPE")) : 1;   // experiments
    WalkArgs walk;
    walk.njobs = 0; walk.first_block[0] = 0;
    for (int cls = 0; cls < 3; cls++) {
        const int cfirst = cls ? list_end[cls - 1] : 0, ccnt = list_end[cls] - cfirst;
        if (ccnt <= 0) continue;
        // lanes per pair and columns per lane: <= 128 | <= 256 | <= 512 target bases
        int seg = cls == 0 ? 16 : 32, maxcpl = cls == 2 ? 16 : 8;
        if (pack_shape == 1) { seg = cls == 0 ? 8 : cls == 1 ? 16 : 32; maxcpl = 16; }
        const int steps_cap = pack_max2[cls] + seg - 1;
        const int per_pass = 2 * (32 / seg);
        const size_t pass_bytes = pack_dirs_bytes(per_pass, seg, maxcpl, steps_cap);
        const long long max_pairs = std::max<long long>(per_pass, (long long)((2ull << 30) / pass_bytes) * per_pass);
        for (long long done = 0; done < ccnt; done += max_pairs) {
            const int cnt = (int)std::min<long long>(max_pairs, ccnt - done);
            const long long npass = ((long long)cnt + per_pass - 1) / per_pass;
            indelgpu_ctx::PackSlot& k = c->pk[walk.njobs];
            if (k.dirs.ensure(pack_dirs_bytes(cnt, seg, maxcpl, steps_cap)) || k.cpl.ensure(4 * (size_t)npass) || k.best.ensure(4 * (size_t)cnt)) {
                cudaStreamSynchronize(st); return INDELGPU_ENOMEM;
            }
            PackArgs w;
            w.n = cnt;
            w.order = c->s_ord.as<int32_t>() + cfirst + done;
            w.targets = c->s_tgt.as<uint8_t>(); w.target_off = c->s_toff.as<int64_t>();
            w.queries = c->s_qry.as<uint8_t>(); w.query_off = c->s_qoff.as<int64_t>();
            w.subs = d_subs; w.indels = d_indels; w.aligned = d_aligned;
            w.dirs = k.dirs.as<uint32_t>(); w.pass_cpl = k.cpl.as<int32_t>(); w.best = k.best.as<uint32_t>();
            w.steps_cap = steps_cap; w.seg = seg; w.maxcpl = maxcpl; w.one = 1u;
            w.error_flag = ctr_err(c);
            int occ = 0;
#define PACK_LAUNCH(S, M) { CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, indel_support_pack_kernel<S, M>, 128, 0)); \
                            if (occ >= 1) indel_support_pack_kernel<S, M><<<(int)std::min<long long>((long long)c->sms * occ, (npass + 3) / 4), 128, 0, st>>>(w); }
            if (seg == 8) PACK_LAUNCH(8, 16)
            else if (seg == 16 && maxcpl == 8) PACK_LAUNCH(16, 8)
            else if (seg == 16) PACK_LAUNCH(16, 16)
            else if (maxcpl == 8) PACK_LAUNCH(32, 8)
            else PACK_LAUNCH(32, 16)
#undef PACK_LAUNCH
            if (occ < 1) { cudaStreamSynchronize(st); return fail(INDELGPU_ELIMIT, "support kernel does not fit on an SM"); }
            c->launches++;
            CU(cudaGetLastError());
            walk.first_block[walk.njobs + 1] = walk.first_block[walk.njobs] + (cnt + 127) / 128;
            walk.job[walk.njobs++] = w;
            if (walk.njobs == kWalkJobs) {                  // every scratch slot is in use
                indel_support_walk_kernel<<<walk.first_block[walk.njobs], 128, 0, st>>>(walk);
                c->launches++;
                CU(cudaGetLastError());
                walk.njobs = 0;
            }
        }
    }
    if (walk.njobs > 0) {
        indel_support_walk_kernel<<<walk.first_block[walk.njobs], 128, 0, st>>>(walk);
        c->launches++;
        CU(cudaGetLastError());
    }
    for (int cls = 0; cls < 3; cls++) {                            // queries too long for it: the carried-counter wavefront
        const int first = list_end[2 + cls], cnt = list_end[3 + cls] - first;
        if (cnt <= 0) continue;
        WaveArgs w;
        w.n = cnt;
        w.order = c->s_ord.as<int32_t>() + first;
        w.targets = c->s_tgt.as<uint8_t>(); w.target_off = c->s_toff.as<int64_t>();
        w.queries = c->s_qry.as<uint8_t>(); w.query_off = c->s_qoff.as<int64_t>();
        w.subs = d_subs; w.indels = d_indels; w.aligned = d_aligned;
        const int group = cls == 0 ? 4 : cls == 1 ? 2 : 1;                  // pairs per warp
        int occ = 0;
        if (cls == 0) CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, indel_support_wave_kernel<8>, 128, 0));
        else if (cls == 1) CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, indel_support_wave_kernel<16>, 128, 0));
        else CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, indel_support_wave_kernel<32>, 128, 0));
        if (occ < 1) { cudaStreamSynchronize(st); return fail(INDELGPU_ELIMIT, "support kernel does not fit on an SM"); }
        const int blocks = (int)std::min<long long>((long long)c->sms * occ, ((long long)cnt + 4 * group - 1) / (4 * group));
        if (cls == 0) indel_support_wave_kernel<8><<<blocks, 128, 0, st>>>(w);
        else if (cls == 1) indel_support_wave_kernel<16><<<blocks, 128, 0, st>>>(w);
        else indel_support_wave_kernel<32><<<blocks, 128, 0, st>>>(w);
        c->launches++;
        CU(cudaGetLastError());
    }
    if (nslow > 0) {
        int occ = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, indel_support_kernel, 128, 0));
        if (occ < 1) { cudaStreamSynchronize(st); return fail(INDELGPU_ELIMIT, "support kernel does not fit on an SM"); }
        const long long cells_cap = (long long)(max1 + 1) * (max2 + 1);
        // resident threads are bounded by the scratch they need (3 bytes per DP cell per thread): at most ~8 GB
        long long blocks = std::min<long long>((long long)c->sms * std::min(occ, 4), (nslow + 127) / 128);
        const long long per_block = 128 * (cells_cap * 3 + 4LL * (max1 + 1));
        blocks = std::max<long long>(1, std::min<long long>(blocks, (8LL << 30) / std::max<long long>(per_block, 1)));
        if (c->s_idx.ensure(4 * (size_t)nslow) ||
            c->s_V.ensure((size_t)blocks * 128 * (size_t)cells_cap * 2) || c->s_I.ensure((size_t)blocks * 128 * (size_t)cells_cap) ||
            c->s_F.ensure((size_t)blocks * 128 * 4 * (size_t)(max1 + 1))) { cudaStreamSynchronize(st); return INDELGPU_ENOMEM; }
        CU(cudaMemcpyAsync(c->s_idx.p, slow.data(), 4 * (size_t)nslow, cudaMemcpyHostToDevice, st));
        SupportArgs a;
        a.n = nslow;
        a.index = c->s_idx.as<int32_t>();
        a.targets = c->s_tgt.as<uint8_t>(); a.target_off = c->s_toff.as<int64_t>();
        a.queries = c->s_qry.as<uint8_t>(); a.query_off = c->s_qoff.as<int64_t>();
        a.subs = d_subs; a.indels = d_indels; a.aligned = d_aligned;
        a.V = c->s_V.as<int16_t>(); a.I = c->s_I.as<int8_t>(); a.F = c->s_F.as<int32_t>();
        a.cells_cap = cells_cap; a.max_len1 = max1; a.max_len2 = max2;
        a.cell_total = ctr_cells(c); a.error_flag = ctr_err(c);
        indel_support_kernel<<<(int)blocks, 128, 0, st>>>(a);
        c->launches++;
        CU(cudaGetLastError());
    }
    CU(cudaEventRecord(c->ev_t1, st));
    c->timed = true;
    CU(cudaMemcpyAsync(h_subs, d_subs, 4 * (size_t)n, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h_indels, d_indels, 4 * (size_t)n, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h_aligned, d_aligned, 4 * (size_t)n, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(c->pinned_small, c->counters.p, 64, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    int err; memcpy(&err, (char*)c->pinned_small + 40, 4);
    if (h_cells) *h_cells = cells;
    if (err) return fail(INDELGPU_ELIMIT, "indel_support_batch: a task exceeds a size limit");
    return 0;
}

// ---------------------------------------------------------------------------------------
// reference-prototype entry points (default context, 1-element batches)
// ---------------------------------------------------------------------------------------
static indelgpu_ctx* g_default = nullptr;
static std::mutex g_default_mu;

[[noreturn]] static void die(const char* what)
{
    // the reference's convention: message on stderr, exit(EXIT_FAILURE) (errors.c:15-27)
    fprintf(stderr, "libindelgpu: %s: %s\n", what, g_err);
    exit(EXIT_FAILURE);
}

static indelgpu_ctx* default_ctx()
{
    if (!g_default) {
        const char* dev = getenv("INDELGPU_DEVICE");
        g_default = indelgpu_create(dev ? atoi(dev) : 0, nullptr);   // scratch is sized per call
        if (!g_default) die("cannot create the default GPU context");
    }
    return g_default;
}

static int script_length(const int* S, int M, int N)
{
    int i = 0, j = 0, k = 0;
    while (i < M || j < N) {
        const int op = S[k++];
        if (op == 0) { i++; j++; } else if (op > 0) j += op; else i -= op;
    }
    return k;
}

extern "C" int local_align(char* seq1, const int seq1len, char* seq2, const int seq2len, const int indx1,
                           const int indx2, int* const psi, int* const psj, int* const pei, int* const pej,
                           int* const S)
{
    std::lock_guard<std::mutex> lk(g_default_mu);
    indelgpu_ctx* c = default_ctx();
    if (seq1len <= 0 || seq2len <= 0) { fprintf(stderr, "Assertion failed: strlen(seq) > 0 file localalign.c\n"); exit(EXIT_FAILURE); }
    if (std::max(-seq1len, indx1) > std::min(seq2len, indx2)) {
        // localalign.c:74-77 prints to stdout; stdout is the VCF stream, so stderr here
        fprintf(stderr, "low > up is unacceptable! low:%d up:%d\n", std::max(-seq1len, indx1), std::min(seq2len, indx2));
        exit(1);
    }
    const int64_t roff[2] = {0, seq1len}, woff[2] = {0, seq2len};
    int32_t low = indx1, up = indx2, score = 0, ends[4], ncig = 0;
    const int stride = seq1len + seq2len + 2;
    std::vector<int32_t> script((size_t)stride);
    if (indelgpu_band_align_batch(c, 1, (const uint8_t*)seq1, roff, (const uint8_t*)seq2, woff, &low, &up,
                                  &score, ends, &ncig, nullptr, 0, script.data(), stride, nullptr))
        die("local_align");
    if (score <= 0) return 0;
    *psi = ends[0]; *psj = ends[1]; *pei = ends[2]; *pej = ends[3];
    const int n = script_length(script.data(), ends[2] - ends[0] + 1, ends[3] - ends[1] + 1);
    memcpy(S, script.data(), sizeof(int) * (size_t)n);
    return score;
}

// one (A, B, S) problem on the device; shared by fetch_cigar and ALIGN
static int upload_pair(indelgpu_ctx* c, const uint8_t* A, int M, const uint8_t* B, int N)
{
    if (c->t_reads.ensure((size_t)std::max(M, 0) + 16) || c->t_refs.ensure((size_t)std::max(N, 0) + 64)) return INDELGPU_ENOMEM;
    if (M > 0) CU(cudaMemcpyAsync(c->t_reads.p, A, (size_t)M, cudaMemcpyHostToDevice, c->stream));
    if (N > 0) CU(cudaMemcpyAsync(c->t_refs.p, B, (size_t)N, cudaMemcpyHostToDevice, c->stream));
    return 0;
}

static int fetch_cigar_gpu(indelgpu_ctx* c, const uint8_t* A, const uint8_t* B, int M, int N, const int* S, int nS,
                           int AP, int readlength, std::vector<uint32_t>& cig, int* mm)
{
    CU(cudaSetDevice(c->device));
    int rc = upload_pair(c, A, M, B, N);
    if (rc) return rc;
    const size_t cap = (size_t)std::max(M, 0) + (size_t)std::max(N, 0) + 4;
    if (c->t_script.ensure(4 * (size_t)(nS + 4)) || c->t_cig.ensure(4 * cap) || c->t_ncig.ensure(16)) return INDELGPU_ENOMEM;
    if (nS > 0) CU(cudaMemcpyAsync(c->t_script.p, S, 4 * (size_t)nS, cudaMemcpyHostToDevice, c->stream));
    fetch_cigar_one_kernel<<<1, 32, 0, c->stream>>>(c->t_reads.as<uint8_t>(), c->t_refs.as<uint8_t>(), M, N,
                                                     c->t_script.as<int>(), AP, readlength, c->t_cig.as<uint32_t>(),
                                                     c->t_ncig.as<int>());
    c->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(c->pinned_small, c->t_ncig.p, 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    int meta[2]; memcpy(meta, c->pinned_small, 8);
    cig.resize((size_t)meta[0]);
    if (meta[0] > 0) CU(cudaMemcpy(cig.data(), c->t_cig.p, 4 * (size_t)meta[0], cudaMemcpyDeviceToHost));
    *mm = meta[1];
    return 0;
}

extern "C" int fetch_cigar(char* A, char* B, int M, int N, int* S, int AP, int BP, const int readlength,
                           int* const pnumops, uint32_t** pcigar)
{
    // A[1], B[1] are the first aligned symbols (alignment.c:374).  The run-length encoding runs on the
    // GPU like everything else; only the caller's realloc'ed buffer convention (globalalign.c:474-476)
    // is handled here.
    (void)BP;
    std::lock_guard<std::mutex> lk(g_default_mu);
    indelgpu_ctx* c = default_ctx();
    c->launches = 0;
    const int nS = script_length(S, M, N);
    std::vector<uint32_t> cig; int mm = 0;
    if (fetch_cigar_gpu(c, (const uint8_t*)A + 1, (const uint8_t*)B + 1, M, N, S, nS, AP, readlength, cig, &mm)) die("fetch_cigar");
    const int n = (int)cig.size();
    if (n > 1) {
        *pcigar = (uint32_t*)realloc(*pcigar, sizeof(uint32_t) * (size_t)n);
        if (!*pcigar) { fprintf(stderr, "libindelgpu: realloc failed\n"); exit(2); }
    }
    for (int t = 0; t < n; t++) (*pcigar)[t] = cig[t];
    *pnumops = n;
    return mm;
}

static int global_align_one(indelgpu_ctx* c, const uint8_t* A, const uint8_t* B, int M, int N, int low, int up,
                            int match, int mismatch, int G, int H, int* S, int* score)
{
    CU(cudaSetDevice(c->device));
    int rc = upload_pair(c, A, M, B, N);
    if (rc) return rc;
    const int m = std::max(M, 0), n = std::max(N, 0);
    const int lo = std::min(std::max(-M, low), std::min(N - M, 0)), hi = std::max(std::min(N, up), std::max(N - M, 0));
    const int band = std::max(hi - lo + 1, 1);
    BandScratch scr;
    scr.max_band = band; scr.max_rows = m; scr.stride = band_scratch_ints(band, m);
    if (c->scratch.ensure(4 * (size_t)scr.stride) || c->t_script.ensure(4 * (size_t)(m + n + 8)) || c->t_ncig.ensure(16)) return INDELGPU_ENOMEM;
    scr.base = c->scratch.as<int>();
    DevParams P = c->P;
    P.match = match; P.mismatch = mismatch; P.G = G; P.H = H;
    global_align_one_kernel<<<1, 32, 0, c->stream>>>(P, scr, c->t_reads.as<uint8_t>(), c->t_refs.as<uint8_t>(), M, N, low, up,
                                                      c->t_script.as<int>(), c->t_ncig.as<int>());
    c->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(c->pinned_small, c->t_ncig.p, 12, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    int meta[3]; memcpy(meta, c->pinned_small, 12);
    if (meta[1] > 0) CU(cudaMemcpy(S, c->t_script.p, 4 * (size_t)meta[1], cudaMemcpyDeviceToHost));
    *score = meta[0];
    return 0;
}

extern "C" int ALIGN(char* A, char* B, int M, int N, int low, int up, int W[][128], int G, int H, int* S)
{
    // A, B point one element before the first symbol (globalalign.h:19-28).  The kernels score
    // by byte equality with one match and one mismatch value, which is the only table the
    // reference ever passes (localalign.c:61-67); verify that instead of silently assuming it.
    std::lock_guard<std::mutex> lk(g_default_mu);
    indelgpu_ctx* c = default_ctx();
    const int match = W['A']['A'], mismatch = W['A']['C'];
    for (int a = 0; a < 128; a++)
        for (int b = 0; b < 128; b++)
            if (W[a][b] != (a == b ? match : mismatch)) {
                fprintf(stderr, "libindelgpu: ALIGN supports match/mismatch tables only (W[%d][%d] = %d)\n", a, b, W[a][b]);
                exit(EXIT_FAILURE);
            }
    c->launches = 0;
    int score = 0;
    if (global_align_one(c, (const uint8_t*)A + 1, (const uint8_t*)B + 1, M, N, low, up, match, mismatch, G, H, S, &score)) die("ALIGN");
    return score;
}

// Debug pretty-printer with the output format of globalalign.c:408-457: blocks of at most 50 alignment
// columns, a ruler, the read row, a marker row ('|' match, ' ' mismatch, '-' gap) and the reference row.
// Pure formatting of a script the GPU produced; host only.
namespace {
struct DisplayBlock {
    std::string top, mid, bot;
    void put(char a, char m, char b) { top.push_back(a); mid.push_back(m); bot.push_back(b); }
    void emit(FILE* F, int block_no, int ap, int bp)
    {
        const int w = (int)top.size();
        fprintf(F, "\n%5d ", 50 * block_no);
        for (int t = 10; t <= w; t += 10) fputs("    .    :", F);
        if (w % 10 >= 5) fputs("    .", F);
        fprintf(F, "\n%5d %s\n      %s\n%5d %s\n", ap, top.c_str(), mid.c_str(), bp, bot.c_str());
        top.clear(); mid.clear(); bot.clear();
    }
};
}  // namespace

extern "C" int DISPLAY(FILE* F, char* A, char* B, int M, int N, int* S, int AP, int BP)
{
    DisplayBlock blk;
    int i = 0, j = 0, blocks = 0, ap = AP, bp = BP, k = 0;
    int gap = 0;                                  // remaining symbols of the gap being printed (sign = kind)
    while (i < M || j < N) {
        if (gap == 0 && S[k] == 0) { k++; const char a = A[++i], b = B[++j]; blk.put(a, a == b ? '|' : ' ', b); }
        else {
            if (gap == 0) gap = S[k++];
            if (gap > 0) { blk.put(' ', '-', B[++j]); gap--; }
            else         { blk.put(A[++i], '-', ' '); gap++; }
        }
        if (blk.top.size() >= 50 || (i >= M && j >= N)) {
            blk.emit(F, blocks++, ap, bp);
            ap = AP + i; bp = BP + j;
        }
    }
    return -1;
}

// pj_bamfilt.cu — device side of `bamfilt`: which alignments survive a junction set.
//
// Reference: BamFilter::filter / containsJunctionInSystem / clipMSR (/root/reference/src/bam_filter.cc:75-245).  An unspliced
// read is kept; a spliced read is kept when at least one of its N ops is a junction of the set.  (In HARD / SOFT clip mode
// the reference "clips" the bad junctions of a multiply spliced read, but only in BamAlignment's C++ CIGAR copy —
// setCigarOpAt, bam_alignment.hpp:177-179 — while BamWriter writes the untouched bam1_t: the files are the same in all three
// modes, only the "Modified" counter differs.  Checked against the unmodified reference in tests/test_bamfilt.py.)
//
// k_bf_keep: one thread per record — the CIGAR walk of A2 (lEnd accumulates reference-consuming ops, an N op of length L
// gives the intron [lEnd, lEnd + L - 1]) and a binary search in the set's (tid, start, end)-sorted arrays.
#include "pj_ctx.hpp"

using namespace pjapi;

struct pj_jset {
    int device = 0; int64_t n = 0;
    int32_t *tid = nullptr, *start = nullptr, *end = nullptr;
    cudaStream_t st = nullptr;
    // per-call scratch (grown on demand)
    int32_t *d_tid = nullptr, *d_pos = nullptr; uint32_t *d_coff = nullptr, *d_cig = nullptr; uint8_t *d_keep = nullptr, *d_nn = nullptr;
    size_t cap_rec = 0, cap_cig = 0;
};

namespace {

__global__ void __launch_bounds__(256) k_bf_keep(int64_t n, const int32_t* __restrict__ tid, const int32_t* __restrict__ pos,
                                                  const uint32_t* __restrict__ cigar_off, const uint32_t* __restrict__ cigar,
                                                  int64_t J, const int32_t* __restrict__ jt, const int32_t* __restrict__ js, const int32_t* __restrict__ je,
                                                  uint8_t* __restrict__ keep, uint8_t* __restrict__ n_nops) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int32_t t = tid[r];
    int32_t lEnd = pos[r];
    uint32_t nn = 0; bool hit = false;
    for (uint32_t k = cigar_off[r]; k < cigar_off[r + 1]; k++) {
        const uint32_t w = __ldg(cigar + k), op = w & 15u; const int32_t L = (int32_t)(w >> 4);
        if (op == 3u) {                                                     // N: intron [lEnd, lEnd + L - 1] (bam_filter.cc:85-93)
            nn++;
            if (!hit) {
                const int32_t s = lEnd, e = lEnd + L - 1;
                int64_t lo = 0, hi = J;
                while (lo < hi) {
                    const int64_t m = (lo + hi) >> 1;
                    const int32_t mt = jt[m], ms = js[m], me = je[m];
                    const bool less = mt != t ? mt < t : ms != s ? ms < s : me < e;
                    if (less) lo = m + 1; else hi = m;
                }
                hit = lo < J && jt[lo] == t && js[lo] == s && je[lo] == e;
            }
            // N does not advance lEnd in the reference's walk (only opConsumesReference ops other than N reach the else branch)
        } else if (op == 0u || op == 2u || op == 7u || op == 8u) lEnd += L;
    }
    keep[r] = (nn == 0 || hit) ? 1 : 0;
    if (n_nops) n_nops[r] = (uint8_t)(nn > 255u ? 255u : nn);
}

} // namespace

extern "C" {

int pj_jset_create(int32_t device, int64_t n, const int32_t* tid, const int32_t* start, const int32_t* end, pj_jset** out) {
    if (!out || n < 0 || (n && (!tid || !start || !end))) return fail(nullptr, PJ_EINVAL, "pj_jset_create: bad arguments");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(nullptr, PJ_ECUDA, "pj_jset_create: no CUDA device available; this library has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(nullptr, PJ_EINVAL, "pj_jset_create: device %d out of range", device);
    for (int64_t i = 1; i < n; i++) {
        const bool ok = tid[i - 1] != tid[i] ? tid[i - 1] < tid[i] : start[i - 1] != start[i] ? start[i - 1] < start[i] : end[i - 1] <= end[i];
        if (!ok) return fail(nullptr, PJ_EINVAL, "pj_jset_create: junctions must be sorted by (tid, start, end)");
    }
    pj_ctx* c = nullptr;
    CU(c, cudaSetDevice(device));
    pj_jset* s = new pj_jset(); s->device = device; s->n = n;
    const size_t N = (size_t)std::max<int64_t>(n, 1);
    auto bail = [&](cudaError_t e) { pj_jset_destroy(s); return fail(nullptr, PJ_ECUDA, "pj_jset_create: %s", cudaGetErrorString(e)); };
    cudaError_t e;
    if ((e = cudaStreamCreateWithFlags(&s->st, cudaStreamNonBlocking)) != cudaSuccess) return bail(e);
    if ((e = cudaMalloc(&s->tid, N * 4)) != cudaSuccess || (e = cudaMalloc(&s->start, N * 4)) != cudaSuccess || (e = cudaMalloc(&s->end, N * 4)) != cudaSuccess) return bail(e);
    if (n) {
        if ((e = cudaMemcpy(s->tid, tid, N * 4, cudaMemcpyHostToDevice)) != cudaSuccess || (e = cudaMemcpy(s->start, start, N * 4, cudaMemcpyHostToDevice)) != cudaSuccess ||
            (e = cudaMemcpy(s->end, end, N * 4, cudaMemcpyHostToDevice)) != cudaSuccess) return bail(e);
    }
    *out = s;
    return PJ_OK;
}

void pj_jset_destroy(pj_jset* s) {
    if (!s) return;
    cudaSetDevice(s->device);
    for (void* p : {(void*)s->tid, (void*)s->start, (void*)s->end, (void*)s->d_tid, (void*)s->d_pos, (void*)s->d_coff, (void*)s->d_cig, (void*)s->d_keep, (void*)s->d_nn}) cudaFree(p);
    if (s->st) cudaStreamDestroy(s->st);
    delete s;
}

int pj_jset_filter(pj_jset* s, int64_t n, const int32_t* tid, const int32_t* pos, const uint32_t* cigar_off, const uint32_t* cigar, uint8_t* keep, uint8_t* n_nops) {
    if (!s || n < 0 || (n && (!tid || !pos || !cigar_off || !keep))) return fail(nullptr, PJ_EINVAL, "pj_jset_filter: bad arguments");
    if (n == 0) return PJ_OK;
    pj_ctx* c = nullptr;
    CU(c, cudaSetDevice(s->device));
    const size_t N = (size_t)n, C = (size_t)(cigar_off[n] - cigar_off[0]);
    if (C && !cigar) return fail(nullptr, PJ_EINVAL, "pj_jset_filter: null cigar column");
    if (cigar_off[0] != 0) return fail(nullptr, PJ_EINVAL, "pj_jset_filter: cigar_off must start at 0");
    if (N > s->cap_rec) {
        for (void* p : {(void*)s->d_tid, (void*)s->d_pos, (void*)s->d_coff, (void*)s->d_keep, (void*)s->d_nn}) cudaFree(p);
        s->cap_rec = N + N / 4 + 1024;
        CU(c, cudaMalloc(&s->d_tid, s->cap_rec * 4)); CU(c, cudaMalloc(&s->d_pos, s->cap_rec * 4)); CU(c, cudaMalloc(&s->d_coff, (s->cap_rec + 1) * 4));
        CU(c, cudaMalloc(&s->d_keep, s->cap_rec)); CU(c, cudaMalloc(&s->d_nn, s->cap_rec));
    }
    if (C > s->cap_cig) { cudaFree(s->d_cig); s->cap_cig = C + C / 4 + 1024; CU(c, cudaMalloc(&s->d_cig, s->cap_cig * 4)); }
    CU(c, cudaMemcpyAsync(s->d_tid, tid, N * 4, cudaMemcpyHostToDevice, s->st)); CU(c, cudaMemcpyAsync(s->d_pos, pos, N * 4, cudaMemcpyHostToDevice, s->st));
    CU(c, cudaMemcpyAsync(s->d_coff, cigar_off, (N + 1) * 4, cudaMemcpyHostToDevice, s->st));
    if (C) CU(c, cudaMemcpyAsync(s->d_cig, cigar, C * 4, cudaMemcpyHostToDevice, s->st));
    k_bf_keep<<<(unsigned)((N + 255) / 256), 256, 0, s->st>>>(n, s->d_tid, s->d_pos, s->d_coff, s->d_cig, s->n, s->tid, s->start, s->end, s->d_keep, s->d_nn);
    CU(c, cudaMemcpyAsync(keep, s->d_keep, N, cudaMemcpyDeviceToHost, s->st));
    if (n_nops) CU(c, cudaMemcpyAsync(n_nops, s->d_nn, N, cudaMemcpyDeviceToHost, s->st));
    CU(c, cudaStreamSynchronize(s->st)); CU(c, cudaGetLastError());
    return PJ_OK;
}

} // extern "C"

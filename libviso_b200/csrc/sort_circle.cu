/*
 * sort_circle.cu -- Match compaction in the reference's std::sort order (viso.cpp:711-724) fused with collect_matches
 * and triangulation, and match_circle (viso.cpp:206-243), with their launch wrappers.
 */
#include "viso_dev.h"
#include "common.cuh"
#include "introsort.h"

/* ------------------------------------------------------------------------------------------------ sort */

/*
 * std::sort order, in parallel.  The reference sorts the Match vector with libstdc++'s UNSTABLE std::sort
 * (viso.cpp:724); which of several equal distances comes first is a property of that algorithm, so the device
 * reproduces the algorithm's data movement exactly instead of using its own sort (introsort.h is the sequential
 * restatement, checked against the real std::sort on the CPU).  Two observations make it parallel:
 *
 *  (1) __unguarded_partition(first+1, last, pivot) swaps the k-th element >= pivot from the left (position L_k) with
 *      the k-th element <= pivot from the right (position R_k) for k = 1..K, K = #{k : L_k < R_k}, and returns
 *      cut = min(L_{K+1}, R_K): both scans only ever read positions the swaps have not touched yet, so the pairing
 *      is a function of the ORIGINAL values.  A warp computes the two position lists with ballots, K with one
 *      monotone predicate, and does all swaps at once.
 *  (2) __final_insertion_sort never moves an element across the boundary of a final partition piece (everything to
 *      the left is <=), and inside a piece it is a stable insertion sort.  So once the <=16-element pieces are known
 *      every element computes its stable rank inside its piece, all in parallel.
 *
 * The median-of-3 pivot moves (3 reads, 1 swap) stay with lane 0; the depth-limit heapsort branch
 * (__partial_sort, taken only by adversarial inputs) is run sequentially by lane 0 with the restated code.
 * Segments are disjoint, so processing the right-hand pieces from a stack instead of by recursion does not change
 * the result.
 */
struct SortScratch {
    int lock, sp, busy; /* work stack of cta_introsort_loop: spin lock, entries, warps that hold a piece */
};

/* the work stack holds disjoint pieces of more than 16 elements: never more than n / 17 of them */
__host__ __device__ inline int sort_stack_entries(int cap) { return cap / 17 + 8; }

/* partition [first+1, last) around the pivot at p[first]; returns cut.  posL / posR: scratch, last-first entries */
__device__ __forceinline__ int warp_partition(viso_sort::KV* p, int first, int last, unsigned short* posL,
                                              unsigned short* posR, int lane)
{
    const int pv = p[first].d;
    int cl = 0, cr = 0;
    for (int base = first + 1; base < last; base += 32) {
        const int i = base + lane;
        const bool in = i < last;
        const int d = in ? p[i].d : 0;
        const bool fL = in && !(d < pv), fR = in && !(pv < d);
        const unsigned mL = __ballot_sync(FULL, fL), mR = __ballot_sync(FULL, fR);
        const unsigned lt = (1u << lane) - 1;
        if (fL) posL[cl + __popc(mL & lt)] = (unsigned short)(i - first);
        if (fR) posR[cr + __popc(mR & lt)] = (unsigned short)(i - first);
        cl += __popc(mL);
        cr += __popc(mR);
    }
    __syncwarp();
    /* R_k (k-th from the right) = posR[cr - k]; L_k = posL[k - 1] */
    const int kmax = min(cl, cr);
    int K = 0;
    for (int k0 = 0; k0 < kmax; k0 += 32) {
        const int k = k0 + lane;
        const bool ok = k < kmax && posL[k] < posR[cr - 1 - k];
        const unsigned m = __ballot_sync(FULL, ok);
        K += __popc(m);
        if (m != FULL) break; /* monotone: the first failure ends it */
    }
    for (int k = lane; k < K; k += 32) {
        const int a = first + posL[k], b = first + posR[cr - 1 - k];
        const viso_sort::KV t = p[a]; p[a] = p[b]; p[b] = t;
    }
    int cut = INT_MAX;
    if (K < cl) cut = first + posL[K];
    if (K > 0) cut = min(cut, first + (int)posR[cr - K]);
    __syncwarp();
    return cut;
}

/* __introsort_loop for p[0..n) by all warps of the CTA; marks the first position of every final piece in leaf[] (pieces
 * sorted by the heapsort branch are marked element by element: they are already in order).
 *
 * libstdc++'s loop partitions a piece, recurses into the right part and goes on with the left; the two parts never
 * interact again and the final arrangement of a piece depends on nothing but its own elements and depth budget, so the
 * recursion is a pool of independent pieces: a warp takes one from a shared stack, partitions it (warp_partition),
 * pushes the right part and keeps the left.  Which warp handles which piece, and when, does not change a single
 * comparison.  One warp used to do all of it (0.18 of the kernel's 0.28 ms per 1000 frames). */
__device__ void cta_introsort_loop(viso_sort::KV* p, int n, unsigned short* posL, unsigned short* posR,
                                   unsigned char* leaf, SortScratch& sc, unsigned short* stk /* [3][entries] */,
                                   int entries, int lane)
{
    volatile int* vlock = &sc.lock;
    volatile int* vsp = &sc.sp;
    volatile int* vbusy = &sc.busy;
    volatile unsigned short* vstk = stk;
    auto acquire = [&]() { while (atomicCAS(&sc.lock, 0, 1) != 0) {} __threadfence_block(); };
    auto release = [&]() { __threadfence_block(); *vlock = 0; };
    for (;;) {
        int got = 0, first = 0, last = 0, depth = 0;
        if (lane == 0) {
            acquire();
            const int sp = *vsp;
            if (sp > 0) {
                first = vstk[sp - 1]; last = vstk[entries + sp - 1]; depth = vstk[2 * entries + sp - 1];
                *vsp = sp - 1;
                *vbusy = *vbusy + 1;
                got = 1;
            } else if (*vbusy == 0) got = -1; /* nothing queued, nobody working: done */
            release();
        }
        got = __shfl_sync(FULL, got, 0);
        if (got < 0) break;
        if (got == 0) { __nanosleep(200); continue; }
        first = __shfl_sync(FULL, first, 0); last = __shfl_sync(FULL, last, 0); depth = __shfl_sync(FULL, depth, 0);
        bool heap_done = false;
        while (last - first > 16) {
            if (depth == 0) {
                if (lane == 0) viso_sort::heap_sort_(p + first, last - first);
                for (int i = first + lane; i < last; i += 32) leaf[i] = 1;
                __syncwarp();
                heap_done = true;
                break;
            }
            --depth;
            const int mid = first + (last - first) / 2;
            if (lane == 0) viso_sort::median_to_first_(p, first, first + 1, mid, last - 1);
            __syncwarp();
            const int cut = warp_partition(p, first, last, posL + first, posR + first, lane);
            if (last - cut > 16) { /* the right part goes on the stack (the partition's writes before the release) */
                if (lane == 0) {
                    acquire();
                    const int sp = *vsp;
                    vstk[sp] = (unsigned short)cut; vstk[entries + sp] = (unsigned short)last; vstk[2 * entries + sp] = (unsigned short)depth;
                    *vsp = sp + 1;
                    release();
                }
            } else if (last > cut && lane == 0) leaf[cut] = 1; /* already a final piece */
            __syncwarp();
            last = cut;
        }
        if (lane == 0) {
            if (!heap_done && last > first) leaf[first] = 1;
            acquire();
            *vbusy = *vbusy - 1;
            release();
        }
        __syncwarp();
    }
}

/*
 * Per frame: (1) compaction of the valid dense results in query order into Match(i, best_idx, best_d1)
 * (viso.cpp:711-722), (2) the reference's std::sort order (viso.cpp:724), (3) pos_of_query inverse table,
 * collect_matches (viso.cpp:501-514) and triangulate_rectified<double> (viso.cpp:1146-1152).
 *
 * When the frame's matches fit the CTA's shared memory (smem_cap of them, 13 bytes each) 8-byte (dist, query)
 * records are sorted there: warp 0 runs the parallel introsort loop, then every thread places one element with its
 * stable rank inside its final piece.  The algorithm only looks at dist, so the resulting permutation is the one
 * std::sort gives the Match vector.  Larger inputs are sorted in place in global memory by one thread with the
 * sequential restatement.
 */
#ifndef VISO_SORT_THREADS
#define VISO_SORT_THREADS 128
#endif
__global__ void __launch_bounds__(VISO_SORT_THREADS) compact_sort_kernel(const SortJob* __restrict__ jobs, ParamDev P, int smem_cap)
{
    extern __shared__ int sort_sm[];
    __shared__ int warp_tot[32];
    __shared__ SortScratch sc;
    viso_sort::KV* kv = reinterpret_cast<viso_sort::KV*>(sort_sm);                    /* [smem_cap] */
    unsigned short* posL = reinterpret_cast<unsigned short*>(sort_sm + 2 * smem_cap); /* [smem_cap] */
    unsigned short* posR = posL + smem_cap;                                            /* [smem_cap] */
    unsigned char* leaf = reinterpret_cast<unsigned char*>(posR + smem_cap);           /* [smem_cap] */
    const int stk_entries = sort_stack_entries(smem_cap);
    unsigned short* stk = reinterpret_cast<unsigned short*>(leaf + ((smem_cap + 15) & ~15)); /* [3][stk_entries] */
    const SortJob job = jobs[blockIdx.x];
    const int n = *job.n;
    int base = 0;
    for (int start = 0; start < n; start += blockDim.x) {
        const int i = start + threadIdx.x;
        int4 r = make_int4(0, 0, 0, 0);
        if (i < n) {
            r = job.dense[i];
            if (job.pos_of_query) job.pos_of_query[i] = -1;
        }
        const bool flag = i < n && r.w != 0;
        const int slot = block_compact_slot(flag, base, warp_tot);
        if (flag) {
            if (slot < smem_cap) { kv[slot].d = r.y; kv[slot].pos = i; leaf[slot] = 0; }
            job.matches[3 * slot + 0] = i;
            job.matches[3 * slot + 1] = r.x;
            job.matches[3 * slot + 2] = r.y;
        }
    }
    const int M = base;
    const bool in_smem = M <= smem_cap;
    if (threadIdx.x == 0) {
        sc.lock = 0; sc.busy = 0; sc.sp = 0;
        if (in_smem && M > 16) { /* the whole array is the first piece */
            stk[0] = 0; stk[stk_entries] = (unsigned short)M; stk[2 * stk_entries] = (unsigned short)(viso_sort::lg(M) * 2);
            sc.sp = 1;
        }
    }
    __syncthreads();
    if (in_smem) {
        if (M > 16) cta_introsort_loop(kv, M, posL, posR, leaf, sc, stk, stk_entries, threadIdx.x & 31);
        else if (M > 0 && threadIdx.x == 0) leaf[0] = 1;
    } else if (threadIdx.x == 0) {
        viso_sort::sort(reinterpret_cast<viso_sort::M3*>(job.matches), M);
    }
    if (threadIdx.x == 0) *job.count = M;
    __syncthreads();
    for (int e = threadIdx.x; e < M; e += blockDim.x) {
        int p, i1, i2, d;
        if (in_smem) {
            /* __final_insertion_sort: stable rank inside the final piece [lo, hi) */
            int lo = e, hi = e + 1;
            while (!leaf[lo]) --lo;
            while (hi < M && !leaf[hi]) ++hi;
            const viso_sort::KV me = kv[e];
            int rank = 0;
            for (int j = lo; j < hi; ++j) {
                const int dj = kv[j].d;
                rank += (dj < me.d || (dj == me.d && j < e)) ? 1 : 0;
            }
            p = lo + rank;
            i1 = me.pos; d = me.d;
            i2 = job.dense[i1].x;
        } else {
            p = e;
            i1 = job.matches[3 * p]; i2 = job.matches[3 * p + 1]; d = job.matches[3 * p + 2];
        }
        if (in_smem) { job.matches[3 * p] = i1; job.matches[3 * p + 1] = i2; job.matches[3 * p + 2] = d; }
        if (job.pos_of_query) job.pos_of_query[i1] = p;
        if (job.x) {
            const float2 a = job.kp1[i1], b = job.kp2[i2];
            const double u1 = a.x, v1 = a.y, u2 = b.x, v2 = b.y;
            const int S = job.stride;
            job.x[0 * S + p] = u1; job.x[1 * S + p] = v1; job.x[2 * S + p] = u2; job.x[3 * S + p] = v2;
            if (job.X) {
                const double dd = u1 - u2;
                job.X[0 * S + p] = P.base * (u1 - P.cu) / dd;
                job.X[1 * S + p] = P.base * (v1 - P.cv) / dd;
                job.X[2 * S + p] = P.f * P.base / dd;
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------------ circle */

/*
 * match_circle, viso.cpp:206-243, for match lists produced by match_desc (unique query index per list): the four
 * nested scans collapse to table lookups -- match11 and match22 are read from the dense per-query results, the
 * position k in match_lr_prev from pos_of_query of the previous frame.  Output order = ascending position i in
 * match_lr, as in the reference.  Also gathers x_c / Xp_c (viso.cpp:1291-1305).
 */
#ifndef VISO_CIRCLE_THREADS
#define VISO_CIRCLE_THREADS 512 /* fewer sequential rounds of the five-deep dependent look-up chain per frame pair (256: +0.03 ms) */
#endif
__global__ void __launch_bounds__(VISO_CIRCLE_THREADS) circle_kernel(const CircleJob* __restrict__ jobs)
{
    __shared__ int warp_tot[32];
    const CircleJob job = jobs[blockIdx.x];
    const int M = *job.lr_count, Mp = *job.lrp_count, npl = *job.n_prev_left;
    const int S = job.stride;
    int base = 0;
    for (int start = 0; start < M; start += blockDim.x) {
        const int i = start + threadIdx.x;
        bool flag = false;
        int il = 0, ir = 0, ilp = 0, irp = 0, k = 0;
        if (i < M) {
            il = job.lr[3 * i]; ir = job.lr[3 * i + 1];
            const int4 a = job.m11[il];
            if (a.w) {
                ilp = a.x;
                if (ilp >= 0 && ilp < npl) {
                    k = job.pos_prev[ilp];
                    if (k >= 0 && k < Mp) {
                        irp = job.lrp[3 * k + 1];
                        const int4 b = job.m22[ir];
                        flag = b.w && b.x == irp;
                    }
                }
            }
        }
        const int c = block_compact_slot(flag, base, warp_tot);
        if (flag) {
            job.circ4[4 * c] = il; job.circ4[4 * c + 1] = ir; job.circ4[4 * c + 2] = ilp; job.circ4[4 * c + 3] = irp;
            job.pcl2[2 * c] = i; job.pcl2[2 * c + 1] = k;
#pragma unroll
            for (int r = 0; r < 4; ++r) job.x_c[r * S + c] = job.x[r * S + i];
#pragma unroll
            for (int r = 0; r < 3; ++r) job.Xp_c[r * S + c] = job.Xp[r * S + k];
        }
    }
    if (threadIdx.x == 0) {
        *job.n_circ = base;
        viso_record_dev rec;
        for (int j = 0; j < 6; ++j) rec.tr[j] = 0;
        rec.ok = 0; rec.n_inliers = 0; rec.n_circ = base; rec.best_hyp = -1;
        *job.rec = rec;
    }
}

/* generic (standalone) variant working from explicit lookup tables, for viso_match_circle() */
__global__ void circle_tables_kernel(const int* __restrict__ m, int n, int* table, int table_n, int key_col, int val_mode,
                                     int* err)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int key = m[3 * i + key_col];
    if (key < 0 || key >= table_n) return;
    const int val = val_mode ? i : m[3 * i + 1];
    const int old = atomicCAS(&table[key], -1, val);
    if (old != -1) atomicOr(err, 2);
}

__global__ void __launch_bounds__(256)
circle_generic_kernel(const int* __restrict__ lr, int nlr, const int* __restrict__ lrp, int nlrp,
                      const int* __restrict__ t11, int n_t11, const int* __restrict__ tlrp, int n_tlrp,
                      const int* __restrict__ t22, int n_t22, int* circ4, int* pcl3, int* n_out)
{
    __shared__ int warp_tot[32];
    int base = 0;
    for (int start = 0; start < nlr; start += blockDim.x) {
        const int i = start + threadIdx.x;
        bool flag = false;
        int il = 0, ir = 0, ilp = 0, irp = 0, k = 0;
        if (i < nlr) {
            il = lr[3 * i]; ir = lr[3 * i + 1];
            if (il >= 0 && il < n_t11 && (ilp = t11[il]) >= 0 && ilp < n_tlrp && (k = tlrp[ilp]) >= 0 && k < nlrp) {
                irp = lrp[3 * k + 1];
                flag = ir >= 0 && ir < n_t22 && t22[ir] == irp && irp >= 0;
            }
        }
        const int c = block_compact_slot(flag, base, warp_tot);
        if (flag) {
            circ4[4 * c] = il; circ4[4 * c + 1] = ir; circ4[4 * c + 2] = ilp; circ4[4 * c + 3] = irp;
            pcl3[3 * c] = i; pcl3[3 * c + 1] = k; pcl3[3 * c + 2] = 0;
        }
    }
    if (threadIdx.x == 0) *n_out = base;
}

/* ------------------------------------------------------------------------------------------------ launchers */

cudaError_t viso_launch_sort(const SortJob* jobs, int n_jobs, int max_n, ParamDev p, cudaStream_t s)
{
    if (n_jobs <= 0) return cudaSuccess;
    /* 13 bytes of shared memory per match: (dist, query) record, two u16 position lists, piece flags */
    int cap = max_n < 1 ? 1 : max_n;
    if (cap > 15000) cap = 15000; /* u16 positions and ~200 KB of shared memory */
    cap = (cap + 3) & ~3;
    const size_t smem = (size_t)cap * 12 + ((cap + 15) & ~15) + 6 * (size_t)sort_stack_entries(cap) + 16;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(compact_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    compact_sort_kernel<<<n_jobs, VISO_SORT_THREADS, smem, s>>>(jobs, p, cap);
    return cudaGetLastError();
}

cudaError_t viso_launch_circle(const CircleJob* jobs, int n_jobs, cudaStream_t s)
{
    if (n_jobs <= 0) return cudaSuccess;
    circle_kernel<<<n_jobs, VISO_CIRCLE_THREADS, 0, s>>>(jobs);
    return cudaGetLastError();
}

cudaError_t viso_launch_circle_tables(const int* m, int n, int* table, int table_n, int key_col, int val_mode,
                                      int* err_flag, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    circle_tables_kernel<<<(n + 255) / 256, 256, 0, s>>>(m, n, table, table_n, key_col, val_mode, err_flag);
    return cudaGetLastError();
}

cudaError_t viso_launch_circle_generic(const int* lr, int nlr, const int* lrp, int nlrp, const int* t11, int n_t11,
                                       const int* tlrp, int n_tlrp, const int* t22, int n_t22,
                                       int* circ4, int* pcl3, int* n_out, cudaStream_t s)
{
    circle_generic_kernel<<<1, 256, 0, s>>>(lr, nlr, lrp, nlrp, t11, n_t11, tlrp, n_tlrp, t22, n_t22, circ4, pcl3, n_out);
    return cudaGetLastError();
}

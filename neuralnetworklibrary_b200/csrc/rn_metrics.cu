// rn_metrics.cu -- the precision/recall integration of mAP1 (reference Applications/Vision.py:1729-1747) for every
// (IoU threshold, category) pair of a validation set in one launch (second half; the IoU matching
// that produces the is_correct flags is rn_map_match in rn_post.cu).
//
// For one (threshold t, category c) the reference
//   * sorts the category's predictions by (score, is_correct) descending   sorted(zip(Scores, IsCorrect), reverse=True)
//   * takes the running count of correct predictions                        np.cumsum(IsCorrect)
//   * multiplies it by 1/n, n = 1..L (float64)                              running_total_true_pos * np.array([1/n ...])
//   * takes the running maximum from the right                              np.flip(np.maximum.accumulate(np.flip(.)))
//   * sums it over the positions of the correct predictions and divides by the number of ground-truth boxes of c
//                                                                           np.sum(precision_smoothed) / ntrue
// Everything is integer or float64 arithmetic on exactly representable inputs, so the table can be reproduced BIT FOR BIT:
// the products and the quotient are single IEEE operations (__dmul_rn / __ddiv_rn) and the final sum follows NumPy's
// pairwise summation (blocks of 128 with 8 interleaved accumulators, recursive halving above that), see rn_pairwise_sum.
//
// One CTA per (category, threshold).  The sort is a rank-by-counting over the category's predictions (keys staged through
// shared memory in tiles; O(L^2) compares spread over 256 threads -- L is the number of predictions of ONE category, a few
// thousand for a COCO-sized validation set), which needs no scratch ordering and is exact for unique keys; the scans are a
// single thread's sequential pass (L steps) while the other 800 CTAs do theirs.  Work is tiny next to the loss kernels: the
// point of doing it here is one launch and one [T, C] read-back instead of a Python loop over categories x thresholds.
#include "rn_common.cuh"

#define RN_AP_THREADS 256
#define RN_AP_TILE 2048

// numpy/core/src/umath/loops_utils.h.src: pairwise_sum for a contiguous float64 array.
static __device__ double rn_pairwise_sum(const double *a, int n) {
    // iterative form of the recursion  sum(a, n) = sum(a, n2) + sum(a + n2, n - n2),  n2 = n/2 rounded down to a multiple of 8
    struct Frame {
        const double *a;
        int n;
        int state;  // 0: enter, 1: left done (value in `left`), 2: both done
        double left;
    };
    Frame st[40];
    int sp = 0;
    st[0].a = a; st[0].n = n; st[0].state = 0; st[0].left = 0.0;
    double ret = 0.0;
    while (sp >= 0) {
        Frame &f = st[sp];
        if (f.state == 0) {
            if (f.n < 8) {
                double res = 0.0;
                for (int i = 0; i < f.n; ++i) res = __dadd_rn(res, f.a[i]);
                ret = res;
                --sp;
            } else if (f.n <= 128) {
                double r[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) r[j] = f.a[j];
                int i;
                for (i = 8; i < f.n - (f.n % 8); i += 8) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], f.a[i + j]);
                }
                double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                                       __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
                for (; i < f.n; ++i) res = __dadd_rn(res, f.a[i]);
                ret = res;
                --sp;
            } else {
                int n2 = f.n / 2;
                n2 -= n2 % 8;
                f.state = 1;
                st[sp + 1].a = f.a; st[sp + 1].n = n2; st[sp + 1].state = 0;
                ++sp;
            }
        } else if (f.state == 1) {
            f.left = ret;
            int n2 = f.n / 2;
            n2 -= n2 % 8;
            f.state = 2;
            st[sp + 1].a = f.a + n2; st[sp + 1].n = f.n - n2; st[sp + 1].state = 0;
            ++sp;
        } else {
            ret = __dadd_rn(f.left, ret);
            --sp;
        }
    }
    return ret;
}

__global__ void __launch_bounds__(RN_AP_THREADS)
rn_map_ap_kernel(const float *__restrict__ scores, const int32_t *__restrict__ perm, const int32_t *__restrict__ cls_off,
                 const unsigned char *__restrict__ is_correct, const int32_t *__restrict__ ntrue, int NP, int C,
                 unsigned long long *__restrict__ keys /*[T][NP]*/, unsigned char *__restrict__ sorted_flag /*[T][NP]*/,
                 double *__restrict__ pv /*[T][NP]*/, double *__restrict__ sm /*[T][NP]*/, double *__restrict__ table /*[T][C]*/) {
    __shared__ unsigned long long s_keys[RN_AP_TILE];
    const int c = blockIdx.x, t = blockIdx.y, tid = threadIdx.x;
    const int seg0 = cls_off[c], L = cls_off[c + 1] - seg0;
    const unsigned char *flag_t = is_correct + (size_t)t * NP;
    unsigned long long *k_seg = keys + (size_t)t * NP + seg0;
    unsigned char *f_seg = sorted_flag + (size_t)t * NP + seg0;
    double *pv_seg = pv + (size_t)t * NP + seg0, *sm_seg = sm + (size_t)t * NP + seg0;

    // keys: (score, is_correct) in the order Python compares the tuples, made unique by the position in the segment
    for (int k = tid; k < L; k += RN_AP_THREADS) {
        const int p = perm[seg0 + k];
        k_seg[k] = ((unsigned long long)rn_float_sortable(scores[p]) << 32) | ((unsigned long long)(flag_t[p] ? 1u : 0u) << 31) |
                   (unsigned long long)(unsigned)k;
    }
    __threadfence_block();
    __syncthreads();
    // rank = number of keys that sort before this one (descending order)
    for (int i0 = 0; i0 < L; i0 += RN_AP_THREADS) {
        const int i = i0 + tid;
        const unsigned long long mine = i < L ? k_seg[i] : 0ull;
        int rank = 0;
        for (int j0 = 0; j0 < L; j0 += RN_AP_TILE) {
            const int nj = min(RN_AP_TILE, L - j0);
            __syncthreads();
            for (int j = tid; j < nj; j += RN_AP_THREADS) s_keys[j] = k_seg[j0 + j];
            __syncthreads();
            if (i < L) {
#pragma unroll 8
                for (int j = 0; j < nj; ++j) rank += (s_keys[j] > mine);
            }
        }
        if (i < L) f_seg[rank] = (unsigned char)((mine >> 31) & 1ull);
    }
    __threadfence_block();
    __syncthreads();
    if (tid != 0) return;
    // the scans (Vision.py:1735-1747), sequential
    long long cum = 0;
    for (int i = 0; i < L; ++i) {
        cum += f_seg[i];
        pv_seg[i] = __dmul_rn((double)cum, __ddiv_rn(1.0, (double)(i + 1)));  // running_total_true_pos * (1/n)
    }
    const int ncorrect = (int)cum;
    double run = -INFINITY;
    int k = ncorrect;
    for (int i = L - 1; i >= 0; --i) {
        run = fmax(run, pv_seg[i]);           // precision_maxes
        if (f_seg[i]) sm_seg[--k] = run;      // precision_smoothed, ascending position order
    }
    const double total = rn_pairwise_sum(sm_seg, ncorrect);
    table[(size_t)t * C + c] = __ddiv_rn(total, (double)ntrue[c]);  // 0/0 = nan for a category without ground truth, like NumPy
}

static inline size_t rn_up256m(size_t x) { return (x + 255) / 256 * 256; }

extern "C" size_t rn_map_ap_workspace_bytes(int NP, int T) {
    if (NP <= 0 || T <= 0) return 256;
    const size_t n = (size_t)NP * (size_t)T;
    return rn_up256m(8 * n) + rn_up256m(n) + 2 * rn_up256m(8 * n);
}

extern "C" int rn_map_ap(const float *pred_scores, const int32_t *perm, const int32_t *cls_off, const unsigned char *is_correct,
                         const int32_t *ntrue, int NP, int C, int T, double *table, void *workspace, size_t workspace_bytes,
                         void *stream) {
    if (NP < 0 || C <= 0 || T <= 0) return rn_set_error(RN_ERR_INVALID_ARG, "rn_map_ap: NP=%d C=%d T=%d", NP, C, T);
    if (C > 65535 || T > 65535) return rn_set_error(RN_ERR_INVALID_ARG, "rn_map_ap: C=%d T=%d too large", C, T);
    if (!cls_off || !ntrue || !table || (NP > 0 && (!pred_scores || !perm || !is_correct)))
        return rn_set_error(RN_ERR_INVALID_ARG, "rn_map_ap: null pointer");
    if (!workspace || workspace_bytes < rn_map_ap_workspace_bytes(NP, T) || (((uintptr_t)workspace) & 255))
        return rn_set_error(RN_ERR_WORKSPACE, "rn_map_ap: workspace needs %zu bytes, 256-byte aligned", rn_map_ap_workspace_bytes(NP, T));
    const size_t n = (size_t)(NP > 0 ? NP : 0) * (size_t)T;
    unsigned char *ws = reinterpret_cast<unsigned char *>(workspace);
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(ws);
    unsigned char *sorted_flag = ws + rn_up256m(8 * n);
    double *pv = reinterpret_cast<double *>(ws + rn_up256m(8 * n) + rn_up256m(n));
    double *sm = reinterpret_cast<double *>(ws + rn_up256m(8 * n) + rn_up256m(n) + rn_up256m(8 * n));
    rn_map_ap_kernel<<<dim3(C, T), RN_AP_THREADS, 0, (cudaStream_t)stream>>>(pred_scores, perm, cls_off, is_correct, ntrue, NP, C,
                                                                            keys, sorted_flag, pv, sm, table);
    return rn_check_launch("rn_map_ap");
}

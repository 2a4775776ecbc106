// rn_abi.cu -- error state, geometry and the small host-only entry points of libretina_sm100.so.
#include <stdarg.h>
#include <stdio.h>
#include <math.h>
#include <string.h>

#include "rn_common.cuh"

static thread_local char g_err[512] = "";

int rn_set_error(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int rn_check_launch(const char *what) {
    cudaError_t e = cudaGetLastError();  // launch-configuration errors only; never synchronises
    if (e != cudaSuccess) return rn_set_error(RN_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
    return RN_OK;
}

extern "C" const char *rn_last_error(void) { return g_err; }
extern "C" int rn_abi_version(void) { return 2; }

// Process-wide tuning / test switches, set explicitly through the ABI (never read from the environment: a stray
// variable in a training job must not change which kernel runs).  0 = default behaviour for every option.
static const char *const g_opt_names[RN_OPT_COUNT] = {"assign_dense", "assign_no_balance", "assign_wbase", "loss_iters",
                                                      "levels_nchunks", "step_fused", "step_bytemap", "loss_prefetch",
                                                      "assign_parts"};
static int g_opt[RN_OPT_COUNT] = {0};

int rn_opt(int id) { return (id >= 0 && id < RN_OPT_COUNT) ? g_opt[id] : 0; }

static int rn_opt_index(const char *name) {
    if (name)
        for (int i = 0; i < RN_OPT_COUNT; ++i)
            if (strcmp(name, g_opt_names[i]) == 0) return i;
    return -1;
}

// The two alternative implementations of rn_loss_step (one persistent kernel; the byte-map chain) measured slower than the
// default and are only compiled with -DRN_EXPERIMENTAL (RN_EXTRA_NVCC_FLAGS at build time): 22 of the library's kernels.
#ifdef RN_EXPERIMENTAL
static const int g_experimental = 1;
#else
static const int g_experimental = 0;
#endif

extern "C" int rn_set_option(const char *name, int value) {
    const int i = rn_opt_index(name);
    if (i < 0) return rn_set_error(RN_ERR_INVALID_ARG, "rn_set_option: unknown option '%s'", name ? name : "(null)");
    if (!g_experimental && value != 0 && (i == RN_OPT_STEP_FUSED || i == RN_OPT_STEP_BYTEMAP))
        return rn_set_error(RN_ERR_INVALID_ARG, "rn_set_option: '%s' needs a library built with -DRN_EXPERIMENTAL", name);
    g_opt[i] = value;
    return RN_OK;
}

extern "C" int rn_get_option(const char *name) {
    if (name && strcmp(name, "experimental") == 0) return g_experimental;  // read-only: how the library was built
    const int i = rn_opt_index(name);
    return i < 0 ? -1 : g_opt[i];
}

extern "C" int rn_num_anchors(int H, int W, int K) {
    if (H <= 0 || W <= 0 || K <= 0) return 0;
    long long cells = 0;
    for (int l = 3; l < 3 + RN_NUM_LEVELS; ++l) {
        long long s = 1LL << l;
        cells += ((H + s - 1) / s) * ((W + s - 1) / s);  // retinanet.py:488
    }
    long long a = cells * K;
    return a > 0x7fffffffLL ? 0 : (int)a;
}

// Fills the geometry either from (H, W, base, K) or, in table mode (anchors != NULL), just A.
int rn_build_geom(RnGeom *g, int H, int W, const double *base, int K, const float *anchors, int A) {
    memset(g, 0, sizeof(*g));
    g->H = H;
    g->W = W;
    g->A = A;
    g->K = K > 0 ? K : 1;
    if (anchors) {
        if (A <= 0) return rn_set_error(RN_ERR_INVALID_ARG, "anchor table given but A=%d", A);
        return RN_OK;
    }
    if (!base) return rn_set_error(RN_ERR_INVALID_ARG, "neither an anchor table nor a base table was given");
    if (K < 1 || K > RN_MAX_K) return rn_set_error(RN_ERR_INVALID_ARG, "K=%d outside [1,%d]", K, RN_MAX_K);
    if (H <= 0 || W <= 0) return rn_set_error(RN_ERR_INVALID_ARG, "bad image size %dx%d", H, W);
    int expect = rn_num_anchors(H, W, K);
    if (expect <= 0 || expect != A)
        return rn_set_error(RN_ERR_INVALID_ARG, "A=%d does not match the %d anchors of a %dx%d image with K=%d", A,
                            expect, H, W, K);
    int off = 0, offc = 0;
    for (int l = 0; l < RN_NUM_LEVELS; ++l) {
        int s = 8 << l;
        g->gh[l] = (H + s - 1) / s;
        g->gw[l] = (W + s - 1) / s;
        g->off[l] = off;
        g->offc[l] = offc;
        off += g->gh[l] * g->gw[l] * K;
        offc += g->gh[l] * g->gw[l];
        double hw = 0.0, hh = 0.0;
        for (int k = 0; k < K; ++k) {
            for (int j = 0; j < 4; ++j) {
                double v = base[(l * K + k) * 4 + j];
                g->base[(l * RN_MAX_K + k) * 4 + j] = v;
                double a = v < 0 ? -v : v;
                if ((j & 1) == 0) hw = a > hw ? a : hw;
                else hh = a > hh ? a : hh;
            }
        }
        g->hw[l] = hw;
        g->hh[l] = hh;
        g->hwf[l] = nextafterf((float)hw, INFINITY);  // >= hw whatever way the cast rounded
        g->hhf[l] = nextafterf((float)hh, INFINITY);
    }
    g->off[RN_NUM_LEVELS] = off;
    g->offc[RN_NUM_LEVELS] = offc;
    return RN_OK;
}

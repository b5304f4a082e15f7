/*
 * capi_internal.h -- definitions shared by the C-ABI translation units (capi.cu: context, standalone entry points,
 * host bookkeeping; capi_seq.cu: the batched sequence pipeline).  Not part of the public ABI.
 */
#ifndef VISO_CAPI_INTERNAL_H_
#define VISO_CAPI_INTERNAL_H_

#include "../../include/viso_b200.h"
#include "viso_dev.h"

#include <string>

/* ------------------------------------------------------------------------------------------------ context */

struct viso_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    int64_t launches = 0;
    GridCfg grid{78, 24}; /* 1248 x 384 px in 16-px cells */
    char* d_scr = nullptr;
    size_t d_cap = 0;
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    cudaStream_t copy_stream = nullptr;   /* host -> device uploads of sequence objects (overlap with compute) */
    cudaStream_t own_copy_stream = nullptr; /* the one this context created (copy_stream may be another context's) */
    int match_mode = VISO_MATCH_AUTO;     /* VISO_MATCH_* (viso_dev.h): which matching kernel path; VISO_MATCH_MODE at viso_create */
    int sm_count = 148;                   /* cudaDevAttrMultiProcessorCount of the device */
    int hyp_it_cap = VISO_HYP_IT_CAP;     /* viso_set_hyp_iteration_cap / VISO_HYP_IT_CAP at viso_create (estimation.cu) */

    int fail(int code, const std::string& msg)
    {
        err = msg;
        return code;
    }
    int fail_cuda(cudaError_t e, const char* what)
    {
        err = std::string(what) + ": " + cudaGetErrorString(e);
        return VISO_ERR_CUDA;
    }
    int ncell() const { return grid.gx * grid.gy; }
};

#define CK(call)                                                        \
    do {                                                                \
        cudaError_t e__ = (call);                                       \
        if (e__ != cudaSuccess) return ctx->fail_cuda(e__, #call);      \
    } while (0)


namespace viso_capi {
ParamDev make_param_dev(const viso_param* p);
MatchParamsDev make_match_dev(const viso_match_params* p);
int status_from_flags(viso_ctx* ctx, int flags);
/* validates the detector geometry and fills the launch configuration; 0 or a VISO_ERR_* with ctx->err set */
int make_harris_cfg(viso_ctx* ctx, int width, int height, int pitch, int n_features, int nbinx, int nbiny, float k,
                    HarrisCfg* out);
} // namespace viso_capi

#endif

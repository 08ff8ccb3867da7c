// dY reconstruction shared by the SIMT (mlp_bwd.cu) and tensor-core (mlp_tc_bwd.cu) backward kernels:
//   dY_l = k1 dz_l - k2 - (y_l - mean) k3r     (GroupNorm backward folded into per-channel coefficients)
// with dz_l either stored (B,C,P) or, for the last layer, synthesised from the pooled gradient and the recorded
// arg-max position.
#pragma once
#include "mlp_common.cuh"

namespace ogc {

struct DySrc {
    int C, P, S, M;
    const float *dz;             // (B,C,P) dense, or NULL -> synthesise from go / sel (last layer)
    const float *go;             // (B,go_ctotal,M) gradient of the pooled output
    const unsigned char *sel;    // (B,C,M) winning position per (channel, centre), 255 = none
    int go_ctotal, go_coff;
    const float *y;              // (B,C,P) pre-norm output of this layer
    const float *coef;           // (B,C,4): k1, k2, k3r, mean
};

struct DyRaw {
    float4 cf;        // k1, k2, k3r, mean
    float yv[4], dzv[4];
};

// Issue the loads for 4 consecutive positions gp..gp+3 (gp % 4 == 0) of channel c (no dependent use).
__device__ __forceinline__ void dy_quad_load(const DySrc &d, int b, int c, int gp, DyRaw &r) {
    r.cf = __ldg(reinterpret_cast<const float4 *>(d.coef + (static_cast<size_t>(b) * d.C + c) * 4));
    const float *yp = d.y + (static_cast<size_t>(b) * d.C + c) * d.P + gp;
    const bool full = gp + 3 < d.P;
    if (full && (reinterpret_cast<uintptr_t>(yp) & 15u) == 0) {
        const float4 t = __ldg(reinterpret_cast<const float4 *>(yp));
        r.yv[0] = t.x; r.yv[1] = t.y; r.yv[2] = t.z; r.yv[3] = t.w;
    } else {
        for (int j = 0; j < 4; ++j) r.yv[j] = gp + j < d.P ? __ldg(yp + j) : 0.f;
    }
    if (d.dz) {
        const float *zp = d.dz + (static_cast<size_t>(b) * d.C + c) * d.P + gp;
        if (full && (reinterpret_cast<uintptr_t>(zp) & 15u) == 0) {
            const float4 t = __ldg(reinterpret_cast<const float4 *>(zp));
            r.dzv[0] = t.x; r.dzv[1] = t.y; r.dzv[2] = t.z; r.dzv[3] = t.w;
        } else {
            for (int j = 0; j < 4; ++j) r.dzv[j] = gp + j < d.P ? __ldg(zp + j) : 0.f;
        }
    } else {
        const int m = gp / d.S, s0 = gp - m * d.S;   // S % 4 == 0: the quad stays inside one centre
        const int sl = gp < d.P ? __ldg(d.sel + (static_cast<size_t>(b) * d.C + c) * d.M + m) : 255;
        const float g = (sl >= s0 && sl < s0 + 4) ? __ldg(d.go + (static_cast<size_t>(b) * d.go_ctotal + d.go_coff + c) * d.M + m) : 0.f;
        for (int j = 0; j < 4; ++j) r.dzv[j] = (sl == s0 + j) ? g : 0.f;
    }
}

__device__ __forceinline__ float4 dy_quad_finish(const DySrc &d, int gp, const DyRaw &r) {
    float o[4];
    for (int j = 0; j < 4; ++j) o[j] = gp + j < d.P ? fmaf(r.cf.x, r.dzv[j], -r.cf.y) - (r.yv[j] - r.cf.w) * r.cf.z : 0.f;
    return make_float4(o[0], o[1], o[2], o[3]);
}

__device__ __forceinline__ float4 dy_quad(const DySrc &d, int b, int c, int gp) {
    DyRaw r;
    dy_quad_load(d, b, c, gp, r);
    return dy_quad_finish(d, gp, r);
}

static inline int fill_dy(DySrc &d, int c, int m, int nsample, const float *dz, const float *go, int go_ctotal, int go_coff,
                   const unsigned char *sel, const float *y, const float *coef) {
    if (!y || !coef) return OGC_ERR_INVALID_ARG;
    if (!dz && (!go || !sel)) return OGC_ERR_INVALID_ARG;
    if (!dz && (nsample % 4 != 0)) return OGC_ERR_UNSUPPORTED;
    d.C = c; d.P = m * nsample; d.S = nsample; d.M = m; d.dz = dz; d.go = go; d.sel = sel;
    d.go_ctotal = go_ctotal; d.go_coff = go_coff; d.y = y; d.coef = coef;
    return OGC_OK;
}


}  // namespace ogc

// g2048_loss.cuh -- per-sample PPO-clip + critic + entropy terms and their analytic gradients
// (train.py:497-554), shared by ppo_loss_kernel (g2048_train.cu) and the fused update kernel
// (g2048_update_x3.cu).
#pragma once
#include <cmath>
#include <cstdint>

namespace g2048 {

// l = logits, m = legal mask, a = action, A = advantage, R = normalised return-to-go:
//   ppo = min(A*rho, A*clamp(rho, 1-eps, 1+eps)), rho = exp(clamp(lp[a] - lp_old, -20, 20))
//   H   = entropy of softmax(clamp(masked l, -20, 20)) over the legal actions;  vl = smooth_l1(V, R)
// Returns ppo, vl, H and  gl[k] = d(ppo + beta_ent*H)/d l_k,  dvl = d vl / d V.
__device__ __forceinline__ void ppo_sample(const float (&l)[4], uint32_t m, uint32_t a, float lp_old, float A, float V,
                                           float R, float clip_eps, float beta_ent, float& ppo_out, float& vl_out,
                                           float& H_out, float (&gl)[4], float& dvl_out) {
    // masked log-softmax
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if ((m >> k) & 1u) mx = fmaxf(mx, l[k]);
    float se = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if ((m >> k) & 1u) se += expf(l[k] - mx);
    const float lse = mx + logf(se);
    float p[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) p[k] = ((m >> k) & 1u) ? expf(l[k] - lse) : 0.f;
    const float la = a == 0 ? l[0] : a == 1 ? l[1] : a == 2 ? l[2] : l[3];
    const float lp_new = la - lse;
    // ratio and clipped surrogate
    const float x = lp_new - lp_old;
    const float xc = fminf(fmaxf(x, -20.f), 20.f);
    const float rho = expf(xc);
    const float rc = fminf(fmaxf(rho, 1.f - clip_eps), 1.f + clip_eps);
    const float t1 = A * rho, t2 = A * rc;
    const float ppo = fminf(t1, t2);
    // d ppo / d rho following torch.minimum (ties split) and clamp (inclusive bounds)
    const bool in_clip = rho >= 1.f - clip_eps && rho <= 1.f + clip_eps;
    float dppo_drho;
    if (t1 < t2) dppo_drho = A;
    else if (t1 > t2) dppo_drho = in_clip ? A : 0.f;
    else dppo_drho = 0.5f * A + (in_clip ? 0.5f * A : 0.f);
    const float dppo_dlp = (x >= -20.f && x <= 20.f) ? dppo_drho * rho : 0.f;
    // entropy over the clamped logits
    float z[4], zmx = -INFINITY;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        z[k] = ((m >> k) & 1u) ? fminf(fmaxf(l[k], -20.f), 20.f) : -20.f;
        zmx = fmaxf(zmx, z[k]);
    }
    float zs = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) zs += expf(z[k] - zmx);
    const float zlse = zmx + logf(zs);
    float q[4], lq[4], H = 0.f, S = 0.f;   // S = sum_{legal} q (1 + log q)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        lq[k] = z[k] - zlse;
        q[k] = expf(lq[k]);
        if ((m >> k) & 1u) {
            H -= q[k] * lq[k];
            S += q[k] * (1.f + lq[k]);
        }
    }
    // critic
    const float d = V - R, ad = fabsf(d);
    const float vl = ad < 1.f ? 0.5f * d * d : ad - 0.5f;
    const float dvl = ad < 1.f ? d : (d > 0.f ? 1.f : -1.f);
    // gradients of loss = -(1/N) sum u
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const bool leg = (m >> k) & 1u;
        const float dlp = leg ? ((k == int(a) ? 1.f : 0.f) - p[k]) : 0.f;          // d lp[a] / d l_k
        const bool pass = leg && l[k] >= -20.f && l[k] <= 20.f;                      // clamp passes gradient
        const float dH = pass ? (-(1.f + lq[k]) * q[k] + q[k] * S) : 0.f;            // d H / d l_k
        gl[k] = dppo_dlp * dlp + beta_ent * dH;
    }
    ppo_out = ppo;
    vl_out = vl;
    H_out = H;
    dvl_out = dvl;
}

}  // namespace g2048

// Coordinate-space kernels of the sampler loop (reference PhysDock/models/model.py:211-281) and the
// RDKit-free part of the physics guidance (template selection :229-243, weighted Kabsch projection
// utils/tensor_utils.py:724-778, direction blend + Euler update model.py:245-250,264,278-281).
// Coordinates reach ~4.6e3 Angstrom in the first steps (1 ulp = 4.9e-4 A), so the elementwise kernels keep the
// reference's operation order with explicit round-to-nearest intrinsics (no FMA contraction).
#include "common.cuh"
#include "kernels.h"

namespace pdk {

namespace {

__device__ __forceinline__ float block_sum(float v, float* scratch) {
    v = warp_sum(v);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    float r = 0.f;
    for (int i = 0; i < nw; ++i) r += scratch[i];
    return r;
}
// One component of _uniform_sphere_point (tensor_utils.py:545-562): phi = u*2*pi, theta = acos(u*2-1); trig evaluated in
// fp64 and rounded once (correctly rounded fp32 results, see glue.cu note).  which: 0 cos(phi), 1 sin(phi), 2 cos(theta),
// 3 sin(theta).  The eight values a rotation needs are evaluated by eight different warps at the same time: a single
// thread doing all ten fp64 transcendentals in sequence was most of the old kernel's 16 us.
__device__ __forceinline__ float sphere_trig(float u_phi, float u_theta, int which) {
    if (which < 2) {
        const float phi = __fmul_rn(__fmul_rn(u_phi, 2.0f), 3.14159265358979323846f);
        return which == 0 ? (float)cos((double)phi) : (float)sin((double)phi);
    }
    const float theta = (float)acos((double)__fsub_rn(__fmul_rn(u_theta, 2.0f), 1.0f));
    return which == 2 ? (float)cos((double)theta) : (float)sin((double)theta);
}

constexpr int COORD_THREADS = 256, COORD_WARPS = COORD_THREADS / 32;
constexpr int COORD_CHUNK = 256;         // atoms transformed per CTA

// centre_random_augmentation (tensor_utils.py:576-586) fused with diffuse (model.py:70-85).
// Grid (atom chunks, samples): every CTA of a sample recomputes the sample's masked mean and rotation (24 KB of
// L2-resident reads and the same arithmetic in the same order, hence bit-identical across CTAs) and transforms its own
// chunk of 256 atoms, so a 2048-atom sample runs on 8 SMs instead of 1.
__global__ void __launch_bounds__(COORD_THREADS) centre_augment_kernel(const float* __restrict__ x, const float* __restrict__ x_exists,
                                                             const float* __restrict__ u4, const float* __restrict__ trans,
                                                             const float* __restrict__ noise, float lambda,
                                                             float noise_scale, float trans_scale,
                                                             float* __restrict__ x_out, int Na) {
    griddep_launch();
    griddep_wait();
    __shared__ float part[COORD_WARPS][4];
    __shared__ float sTrig[8];
    __shared__ float sR[9];
    __shared__ float sMean[3];
    const int b = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float* xb = x + (size_t)b * Na * 3;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, sm = 0.f;
    for (int i = tid; i < Na; i += COORD_THREADS) {
        const float m = x_exists[i];
        s0 += xb[3 * i] * m; s1 += xb[3 * i + 1] * m; s2 += xb[3 * i + 2] * m; sm += m;
    }
    s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2); sm = warp_sum(sm);
    if (lane == 0) {
        part[warp][0] = s0; part[warp][1] = s1; part[warp][2] = s2; part[warp][3] = sm;
        // warp w evaluates trig value w: (cos phi0, sin phi0, cos theta0, sin theta0, cos phi1, sin phi1, cos theta1, sin theta1)
        const int pt = warp >> 2;
        sTrig[warp] = sphere_trig(u4[4 * b + 2 * pt], u4[4 * b + 2 * pt + 1], warp & 3);
    }
    __syncthreads();
    if (tid == 0) {
        float t[4] = {0.f, 0.f, 0.f, 0.f};
        for (int w = 0; w < COORD_WARPS; ++w)
            for (int k = 0; k < 4; ++k) t[k] += part[w][k];
        sMean[0] = t[0] / t[3]; sMean[1] = t[1] / t[3]; sMean[2] = t[2] / t[3];
        float e0[3], u1[3], e1[3];
        e0[0] = __fmul_rn(sTrig[0], sTrig[3]); e0[1] = __fmul_rn(sTrig[1], sTrig[3]); e0[2] = sTrig[2];
        u1[0] = __fmul_rn(sTrig[4], sTrig[7]); u1[1] = __fmul_rn(sTrig[5], sTrig[7]); u1[2] = sTrig[6];
        // uniform_random_rotation (tensor_utils.py:565-573): Gram-Schmidt + cross product, rows e0,e1,e2
        const float dot = __fadd_rn(__fadd_rn(__fmul_rn(u1[0], e0[0]), __fmul_rn(u1[1], e0[1])), __fmul_rn(u1[2], e0[2]));
        for (int k = 0; k < 3; ++k) e1[k] = __fsub_rn(u1[k], __fmul_rn(e0[k], dot));
        const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(e1[0], e1[0]), __fmul_rn(e1[1], e1[1])), __fmul_rn(e1[2], e1[2])));
        for (int k = 0; k < 3; ++k) e1[k] = e1[k] / nrm;
        sR[0] = e0[0]; sR[1] = e0[1]; sR[2] = e0[2];
        sR[3] = e1[0]; sR[4] = e1[1]; sR[5] = e1[2];
        sR[6] = __fsub_rn(__fmul_rn(e0[1], e1[2]), __fmul_rn(e0[2], e1[1]));
        sR[7] = __fsub_rn(__fmul_rn(e0[2], e1[0]), __fmul_rn(e0[0], e1[2]));
        sR[8] = __fsub_rn(__fmul_rn(e0[0], e1[1]), __fmul_rn(e0[1], e1[0]));
    }
    __syncthreads();
    const float t0 = __fmul_rn(trans_scale, trans[3 * b]), t1 = __fmul_rn(trans_scale, trans[3 * b + 1]),
                t2 = __fmul_rn(trans_scale, trans[3 * b + 2]);
    const int i_end = min(Na, (int)(blockIdx.x + 1) * COORD_CHUNK);
    for (int i = blockIdx.x * COORD_CHUNK + tid; i < i_end; i += COORD_THREADS) {
        const float a0 = __fsub_rn(xb[3 * i], sMean[0]), a1 = __fsub_rn(xb[3 * i + 1], sMean[1]),
                    a2 = __fsub_rn(xb[3 * i + 2], sMean[2]);
        float o[3];
#pragma unroll
        for (int r = 0; r < 3; ++r)   // einsum "...ij,...kj->...ki": out_r = sum_j R[r][j] * a_j
            o[r] = fmaf(sR[3 * r + 2], a2, fmaf(sR[3 * r + 1], a1, __fmul_rn(sR[3 * r], a0)));
        o[0] = __fadd_rn(o[0], t0); o[1] = __fadd_rn(o[1], t1); o[2] = __fadd_rn(o[2], t2);
        if (noise != nullptr) {       // x_cur + (lambda * noise) * sqrt(t_hat^2 - t_cur^2)
            const float* nz = noise + ((size_t)b * Na + i) * 3;
#pragma unroll
            for (int r = 0; r < 3; ++r) o[r] = __fadd_rn(o[r], __fmul_rn(__fmul_rn(lambda, nz[r]), noise_scale));
        }
        float* dst = x_out + ((size_t)b * Na + i) * 3;
        dst[0] = o[0]; dst[1] = o[1]; dst[2] = o[2];
    }
}

// d_cur and the Euler update (model.py:245-250,263-264,278-281), reference operation order, no contraction.
__global__ void __launch_bounds__(256) euler_kernel(const float* __restrict__ x_hat, const float* __restrict__ x_den,
                                                    const float* __restrict__ aligned, const float* __restrict__ w,
                                                    const float* __restrict__ t_hat, float t_next, float eta,
                                                    float* __restrict__ x_next, int B, int Na) {
    griddep_launch();
    griddep_wait();
    const size_t total = (size_t)B * Na * 3;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t r = i / 3;
        const int at = (int)(r % Na), b = (int)(r / Na);
        const float th = t_hat[b];
        const float xh = x_hat[i];
        float d = __fsub_rn(xh, x_den[i]) / th;
        if (aligned != nullptr) {
            const float wi = w[at];
            const float dl = __fmul_rn(__fsub_rn(xh, aligned[i]) / th, wi);
            d = __fadd_rn(__fmul_rn(d, __fsub_rn(1.0f, wi)), dl);
        }
        const float step = __fmul_rn(eta, __fsub_rn(t_next, th));
        x_next[i] = __fadd_rn(xh, __fmul_rn(step, d));
    }
}

// epsilon[b,c] of model.py:231-239.  CTA per (template c, sample b).
__global__ void __launch_bounds__(128) template_eps_kernel(const float* __restrict__ x_den, const int* __restrict__ lig_idx,
                                                           const float* __restrict__ ref_dist, float* __restrict__ eps,
                                                           int Na, int n, int C) {
    griddep_launch();
    griddep_wait();
    extern __shared__ float sl[];     // [n][3]
    __shared__ float scratch[4];
    const int c = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    for (int i = tid; i < n * 3; i += blockDim.x) sl[i] = x_den[((size_t)b * Na + lig_idx[i / 3]) * 3 + i % 3];
    __syncthreads();
    const float* rd = ref_dist + (size_t)c * n * n;
    float acc = 0.f;
    for (int q = tid; q < n * n; q += blockDim.x) {
        const int i = q / n, j = q % n;
        const float dx = sl[3 * i] - sl[3 * j], dy = sl[3 * i + 1] - sl[3 * j + 1], dz = sl[3 * i + 2] - sl[3 * j + 2];
        const float dist = sqrtf(dx * dx + dy * dy + dz * dz);
        const float delta = fabsf(dist - rd[q]);
        const float e = 1.f / (1.f + expf(0.5f - delta)) + 1.f / (1.f + expf(1.f - delta)) +
                        1.f / (1.f + expf(2.f - delta)) + 1.f / (1.f + expf(4.f - delta));
        acc += 0.25f * e;
    }
    acc = block_sum(acc, scratch);
    if (tid == 0) eps[(size_t)b * C + c] = acc / (float)(n * n);
}

// argmin over templates (first minimum, like torch.argmin on CPU) + scatter of the chosen template into
// batch_ref_pos[:, is_ligand_atom] (model.py:240-241).  CTA per sample.
__global__ void __launch_bounds__(128) template_pick_kernel(const float* __restrict__ eps, const float* __restrict__ ref_poses,
                                                            const int* __restrict__ lig_idx, int64_t* __restrict__ used,
                                                            float* __restrict__ batch_ref_pos, int Na, int n, int C) {
    griddep_launch();
    griddep_wait();
    __shared__ int best;
    const int b = blockIdx.x;
    if (threadIdx.x == 0) {
        int bi = 0;
        float bv = eps[(size_t)b * C];
        for (int c = 1; c < C; ++c) {
            const float v = eps[(size_t)b * C + c];
            if (v < bv) { bv = v; bi = c; }
        }
        best = bi;
        used[b] = bi;
    }
    __syncthreads();
    const float* src = ref_poses + (size_t)best * n * 3;
    for (int i = threadIdx.x; i < n * 3; i += blockDim.x)
        batch_ref_pos[((size_t)b * Na + lig_idx[i / 3]) * 3 + i % 3] = src[i];
}

// ---- 3x3 Kabsch in fp64 ----------------------------------------------------------------------------
// Q = argmax_{rotation} tr(Q H),  H[j][k] = sum_i w_i g_i[j] p_i[k]  (tensor_utils.py:757-773 computes the same Q
// as transpose(U diag(1,1,det) Vh) from torch.linalg.svd).  Here: Jacobi eigen-decomposition of H^T H = V S^2 V^T,
// u_k = H v_k / s_k, third axes completed by cross products, which yields the reflection-corrected optimum
// directly (det V = det U = +1) and stays well defined for planar ligands (s_3 -> 0).
__device__ void kabsch3(const double H[3][3], double Q[3][3]) {
    double K[3][3], V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) K[i][j] = H[0][i] * H[0][j] + H[1][i] * H[1][j] + H[2][i] * H[2][j];
    for (int sweep = 0; sweep < 30; ++sweep) {
        const double off = fabs(K[0][1]) + fabs(K[0][2]) + fabs(K[1][2]);
        if (off < 1e-300 || off <= 1e-18 * (fabs(K[0][0]) + fabs(K[1][1]) + fabs(K[2][2]))) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (fabs(K[p][q]) < 1e-300) continue;
                const double theta = (K[q][q] - K[p][p]) / (2.0 * K[p][q]);
                const double tt = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double cs = 1.0 / sqrt(tt * tt + 1.0), sn = tt * cs;
                for (int k = 0; k < 3; ++k) {   // K <- K J
                    const double kp = K[k][p], kq = K[k][q];
                    K[k][p] = cs * kp - sn * kq; K[k][q] = sn * kp + cs * kq;
                }
                for (int k = 0; k < 3; ++k) {   // K <- J^T K
                    const double kp = K[p][k], kq = K[q][k];
                    K[p][k] = cs * kp - sn * kq; K[q][k] = sn * kp + cs * kq;
                }
                for (int k = 0; k < 3; ++k) {   // V <- V J
                    const double vp = V[k][p], vq = V[k][q];
                    V[k][p] = cs * vp - sn * vq; V[k][q] = sn * vp + cs * vq;
                }
            }
    }
    int o0 = 0, o1 = 1, o2 = 2;     // order eigenvalues descending
    if (K[o0][o0] < K[o1][o1]) { int t = o0; o0 = o1; o1 = t; }
    if (K[o0][o0] < K[o2][o2]) { int t = o0; o0 = o2; o2 = t; }
    if (K[o1][o1] < K[o2][o2]) { int t = o1; o1 = o2; o2 = t; }
    double v1[3] = {V[0][o0], V[1][o0], V[2][o0]}, v2[3] = {V[0][o1], V[1][o1], V[2][o1]};
    double v3[3] = {v1[1] * v2[2] - v1[2] * v2[1], v1[2] * v2[0] - v1[0] * v2[2], v1[0] * v2[1] - v1[1] * v2[0]};
    double u1[3], u2[3], u3[3];
    for (int i = 0; i < 3; ++i) {
        u1[i] = H[i][0] * v1[0] + H[i][1] * v1[1] + H[i][2] * v1[2];
        u2[i] = H[i][0] * v2[0] + H[i][1] * v2[1] + H[i][2] * v2[2];
    }
    double n1 = sqrt(u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2]);
    if (n1 < 1e-300) {              // H == 0: identity
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) Q[i][j] = (i == j);
        return;
    }
    for (int i = 0; i < 3; ++i) u1[i] /= n1;
    const double d12 = u1[0] * u2[0] + u1[1] * u2[1] + u1[2] * u2[2];
    for (int i = 0; i < 3; ++i) u2[i] -= d12 * u1[i];
    double n2 = sqrt(u2[0] * u2[0] + u2[1] * u2[1] + u2[2] * u2[2]);
    if (n2 < 1e-12 * n1) {          // rank-1 H: any unit vector orthogonal to u1
        const int k = fabs(u1[0]) < fabs(u1[1]) ? (fabs(u1[0]) < fabs(u1[2]) ? 0 : 2) : (fabs(u1[1]) < fabs(u1[2]) ? 1 : 2);
        double e[3] = {0, 0, 0};
        e[k] = 1.0;
        const double de = u1[k];
        for (int i = 0; i < 3; ++i) u2[i] = e[i] - de * u1[i];
        n2 = sqrt(u2[0] * u2[0] + u2[1] * u2[1] + u2[2] * u2[2]);
    }
    for (int i = 0; i < 3; ++i) u2[i] /= n2;
    u3[0] = u1[1] * u2[2] - u1[2] * u2[1]; u3[1] = u1[2] * u2[0] - u1[0] * u2[2]; u3[2] = u1[0] * u2[1] - u1[1] * u2[0];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) Q[i][j] = v1[i] * u1[j] + v2[i] * u2[j] + v3[i] * u3[j];   // Q = V U^T
}

// Sums K per-thread fp64 values over the CTA with ONE barrier pair (the old code called block_sum once per value:
// 16 x 2 barriers per launch); every thread returns with all K totals, summed in a fixed order.
template <int K>
__device__ __forceinline__ void block_sum_n(double (&v)[K], double (*part)[16]) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = warp_sum(v[k]);
    __syncthreads();                  // `part` may still be read from the previous call
    if (lane == 0)
#pragma unroll
        for (int k = 0; k < K; ++k) part[warp][k] = v[k];
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double r = 0.0;
        for (int w = 0; w < COORD_WARPS; ++w) r += part[w][k];
        v[k] = r;
    }
}

// weighted_rigid_align (tensor_utils.py:724-778): returns x_gt rotated + translated onto x_pred's frame.
// x_pred = x_den * x_exists is formed on the fly (model.py:245).  Grid (atom chunks, samples): like centre_augment every
// CTA of a sample recomputes the (tiny, ligand-only) moments and the 3x3 Kabsch and transforms its own 256 atoms.
__global__ void __launch_bounds__(COORD_THREADS) rigid_align_kernel(const float* __restrict__ x_den, const float* __restrict__ x_exists,
                                                          const float* __restrict__ x_gt, int gt_batched,
                                                          const float* __restrict__ w, float* __restrict__ aligned, int Na) {
    griddep_launch();
    griddep_wait();
    __shared__ double part[COORD_WARPS][16];
    __shared__ float sQ[9], sMuP[3], sMuG[3];
    const int b = blockIdx.y, tid = threadIdx.x;
    const float* xp = x_den + (size_t)b * Na * 3;
    const float* xg = x_gt + (gt_batched ? (size_t)b * Na * 3 : 0);
    double acc[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int i = tid; i < Na; i += COORD_THREADS) {
        const float wi = w[i];
        if (wi != 0.f) {
            const float m = x_exists[i];
            acc[0] += wi;
            for (int k = 0; k < 3; ++k) { acc[1 + k] += (double)wi * (xp[3 * i + k] * m); acc[4 + k] += (double)wi * xg[3 * i + k]; }
        }
    }
    block_sum_n<7>(acc, part);
    const float muP[3] = {(float)(acc[1] / acc[0]), (float)(acc[2] / acc[0]), (float)(acc[3] / acc[0])};
    const float muG[3] = {(float)(acc[4] / acc[0]), (float)(acc[5] / acc[0]), (float)(acc[6] / acc[0])};
    double h[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = tid; i < Na; i += COORD_THREADS) {
        const float wi = w[i];
        if (wi != 0.f) {
            const float m = x_exists[i];
            float gp[3], pp[3];
            for (int k = 0; k < 3; ++k) { gp[k] = xg[3 * i + k] - muG[k]; pp[k] = xp[3 * i + k] * m - muP[k]; }
            for (int j = 0; j < 3; ++j)
                for (int k = 0; k < 3; ++k) h[3 * j + k] += (double)wi * gp[j] * pp[k];
        }
    }
    block_sum_n<9>(h, part);
    if (tid == 0) {
        double H[3][3], Q[3][3];
        for (int j = 0; j < 3; ++j)
            for (int k = 0; k < 3; ++k) H[j][k] = h[3 * j + k];
        kabsch3(H, Q);
        for (int j = 0; j < 3; ++j)
            for (int k = 0; k < 3; ++k) sQ[3 * j + k] = (float)Q[j][k];
        for (int k = 0; k < 3; ++k) { sMuP[k] = muP[k]; sMuG[k] = muG[k]; }
    }
    __syncthreads();
    const int i_end = min(Na, (int)(blockIdx.x + 1) * COORD_CHUNK);
    for (int i = blockIdx.x * COORD_CHUNK + tid; i < i_end; i += COORD_THREADS) {
        const float g0 = xg[3 * i] - sMuG[0], g1 = xg[3 * i + 1] - sMuG[1], g2 = xg[3 * i + 2] - sMuG[2];
        float* dst = aligned + ((size_t)b * Na + i) * 3;
#pragma unroll
        for (int r = 0; r < 3; ++r)
            dst[r] = fmaf(sQ[3 * r + 2], g2, fmaf(sQ[3 * r + 1], g1, sQ[3 * r] * g0)) + sMuP[r];
    }
}

inline int grid_for(size_t n, int block = 256, int cap = 148 * 8) {
    size_t g = (n + block - 1) / block;
    return (int)(g < (size_t)cap ? (g ? g : 1) : cap);
}

}  // namespace

cudaError_t launch_centre_augment(const float* x, const float* x_exists, const float* u4, const float* trans,
                                  const float* noise, float lambda, float noise_scale, float trans_scale,
                                  float* x_out, int B, int Na, cudaStream_t st) {
    if (B <= 0 || Na <= 0) return cudaErrorInvalidValue;
    PDK_LAUNCH_CHECK(launch_pdl(centre_augment_kernel, dim3((Na + COORD_CHUNK - 1) / COORD_CHUNK, B), dim3(COORD_THREADS), (size_t)(0), st, x, x_exists, u4, trans, noise, lambda, noise_scale, trans_scale, x_out, Na));
    return cudaGetLastError();
}

cudaError_t launch_euler(const float* x_hat, const float* x_den, const float* aligned, const float* w,
                         const float* t_hat, float t_next, float eta, float* x_next, int B, int Na,
                         cudaStream_t st) {
    if (B <= 0 || Na <= 0 || (aligned != nullptr && w == nullptr)) return cudaErrorInvalidValue;
    PDK_LAUNCH_CHECK(launch_pdl(euler_kernel, dim3(grid_for((size_t)B * Na * 3)), dim3(256), (size_t)(0), st, x_hat, x_den, aligned, w, t_hat, t_next, eta, x_next, B, Na));
    return cudaGetLastError();
}

cudaError_t launch_template_eps(const float* x_den, const int* lig_idx, const float* ref_dist, float* eps,
                                int B, int Na, int n_lig, int C, cudaStream_t st) {
    if (B <= 0 || n_lig <= 0 || C <= 0 || (size_t)n_lig * 12 > 48 * 1024) return cudaErrorInvalidValue;
    PDK_LAUNCH_CHECK(launch_pdl(template_eps_kernel, dim3(C, B), dim3(128), (size_t)n_lig * 3 * sizeof(float), st, x_den, lig_idx, ref_dist, eps, Na, n_lig, C));
    return cudaGetLastError();
}

cudaError_t launch_template_pick(const float* eps, const float* ref_poses, const int* lig_idx, int64_t* used,
                                 float* batch_ref_pos, int B, int Na, int n_lig, int C, cudaStream_t st) {
    if (B <= 0 || n_lig <= 0 || C <= 0) return cudaErrorInvalidValue;
    PDK_LAUNCH_CHECK(launch_pdl(template_pick_kernel, dim3(B), dim3(128), (size_t)(0), st, eps, ref_poses, lig_idx, used, batch_ref_pos, Na, n_lig, C));
    return cudaGetLastError();
}

cudaError_t launch_rigid_align(const float* x_den, const float* x_exists, const float* x_gt, int gt_batched,
                               const float* w, float* aligned, int B, int Na, cudaStream_t st) {
    if (B <= 0 || Na <= 0) return cudaErrorInvalidValue;
    PDK_LAUNCH_CHECK(launch_pdl(rigid_align_kernel, dim3((Na + COORD_CHUNK - 1) / COORD_CHUNK, B), dim3(COORD_THREADS), (size_t)(0), st, x_den, x_exists, x_gt, gt_batched, w, aligned, Na));
    return cudaGetLastError();
}

// Pairwise pose RMSD matrix used for ranking (redocking.py:391:
//   dist = sqrt(mean_atoms(|pred[:,None] - pred[None]|^2))), fp64 accumulation.  One warp per (s, t) pair.
namespace {
__global__ void __launch_bounds__(256) pairwise_rmsd_kernel(const float* __restrict__ poses, double* __restrict__ dist, int S, int n) {
    griddep_launch();
    griddep_wait();
    const int pair = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (pair >= S * S) return;
    const int a = pair / S, b = pair % S;
    const float* pa = poses + (size_t)a * n * 3;
    const float* pb = poses + (size_t)b * n * 3;
    double acc = 0.0;
    for (int i = lane; i < n; i += 32) {
        const double dx = (double)pa[3 * i] - (double)pb[3 * i], dy = (double)pa[3 * i + 1] - (double)pb[3 * i + 1],
                     dz = (double)pa[3 * i + 2] - (double)pb[3 * i + 2];
        const double nrm = sqrt(dx * dx + dy * dy + dz * dz);      // np.linalg.norm(...) ** 2, as the reference writes it
        acc += nrm * nrm;
    }
    acc = warp_sum(acc);
    if (lane == 0) dist[pair] = sqrt(acc / (double)n);
}
}  // namespace

cudaError_t launch_pairwise_rmsd(const float* poses, double* dist, int S, int n, cudaStream_t st) {
    if (S <= 0 || n <= 0) return cudaErrorInvalidValue;
    PDK_LAUNCH_CHECK(launch_pdl(pairwise_rmsd_kernel, dim3((S * S + 7) / 8), dim3(256), (size_t)0, st, poses, dist, S, n));
    return cudaGetLastError();
}

}  // namespace pdk

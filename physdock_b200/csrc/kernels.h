// Internal launcher declarations shared by the kernels and the C ABI (capi.cu).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace pdk {

// ----------------------------------------------------------------------------- split-fp16 GEMM
// C[M,N] = A[M,K] * W[N,K]^T with A, W given as (hi, lo) fp16 planes; fp32 accumulation of
// Ah*Wh + Ah*Wl + Al*Wh.  M % 128 == 0, N % 128 == 0, K % 64 == 0 (activations are padded to S_pad).
enum GemmEpilogue {
    EPI_STORE = 0,       // out[M,N] (fp32) = act(acc + bias)
    EPI_GATE_RESID = 1,  // out[M,N] (fp32, in place) += (acc + bias) * gate[sample(row)][col]
    EPI_SWIGLU = 2,      // W rows interleaved in blocks of 16 (w1 | w3): planes[M, N/2] = split(silu(h1) * h3)
    EPI_QKV = 3,         // N = 3c: per-head RMSNorm on q and k, q pre-scaled, interleaved planes [B,H,S_pad,64]
};

struct GemmArgs {
    const __half* Ah; const __half* Al; int lda;
    const __half* Wh; const __half* Wl; int ldw;
    int M, N, K;
    const float* bias;          // [N] or nullptr
    float* out; int ldo;        // EPI_STORE / EPI_GATE_RESID
    int act_silu;               // EPI_STORE only
    const float* gate;          // EPI_GATE_RESID: gate[sample * gate_stride + col]
    int gate_stride;
    int rows_per_sample;        // S_pad (EPI_GATE_RESID, EPI_QKV)
    __half* ph; __half* pl; int ldp;   // EPI_SWIGLU output planes [M, N/2]
    __half* q; __half* k; __half* v;   // EPI_QKV: [B,H,S_pad,64] rows = [hi 32 | lo 32] (128-byte rows for TMA)
    const float* norm_q; const float* norm_k;   // [32] RMSNorm gains
    int c;                      // model width (N == 3c), a power of two
    int c_shift;                // log2(c), filled in by launch_gemm
    float rms_eps; float q_scale;
};
cudaError_t launch_gemm(GemmEpilogue epi, const GemmArgs& a, cudaStream_t st);

// ----------------------------------------------------------------------------- fused transition (transition_umma.cu), c = 128
// x[M,128] += w2( SiLU(w1 xn) * (w3 xn) ) * gate, xn = LN(x) * (1 + scale) + shift; (shift, scale, gate) = mod[sample, off..off+3c)
struct TransitionArgs {
    float* x;                       // [M,128] fp32, updated in place
    const float* mod; int mod_stride; int mod_off;
    const __half* w13h; const __half* w13l;     // [2*hidden,128], rows interleaved (16 x w1 | 16 x w3)
    const __half* w2h; const __half* w2l;       // [128,hidden]
    int M, hidden, rows_per_sample;
    float eps;
};
cudaError_t launch_transition_fused(const TransitionArgs& a, cudaStream_t st);

// ----------------------------------------------------------------------------- pair-bias attention
// q,k,v [B,H,S_pad,64] with rows [hi 32 | lo 32] (q pre-scaled by log2e/sqrt(32)); bias [H,S_pad,S_pad] fp32
// (pre-scaled by log2e, pad columns = kPadBias); output planes o[B*S_pad, c] with column h*32+d.
struct AttnArgs {
    const __half* q; const __half* k; const __half* v;
    const float* bias;
    __half* oh; __half* ol;
    int B, H, S_pad, c;
    int b_base;         // set by launch_attention: first sample of the launched chunk
    long long* trace;   // debug only: per-unit clock64 stamps of CTA (0,0,0); nullptr in production
};
cudaError_t launch_attention(const AttnArgs& a, cudaStream_t st);
// entries (head | q tile << 8 | first sample << 16 | samples << 24) of the balanced work list; -1: does not fit one launch
int attention_work_list(int B, int H, int QT, int sms, uint32_t* out, int cap);
void set_attention_trace(long long* buf);   // debug hook used by tools/trace_attention.py

// ----------------------------------------------------------------------------- pair-bias prepass
// bias[l][h][i][j] = log2e * ( sum_c wfold[l*H+h][c] * xhat(pair[i][j])[c] + bfold[l*H+h] + (mask==0 ? -inf_ : 0) )
// for i,j < S; pad columns kPadBias; pad rows 0.   pair [S,S,C] fp32, C in {16,128}, LH = L*H outputs per pair.
cudaError_t launch_pair_bias(const float* pair, const float* mask, const float* wfold, const float* bfold,
                             float* bias, int S, int S_pad, int C, int LH, float ln_eps, float inf_,
                             cudaStream_t st);

// ----------------------------------------------------------------------------- conditioning / glue
// t_hat[B] -> SiLU(time_embedder(t_hat * c_noise)) as fp32 tsilu[B,256] (optional) and/or as split planes
// ts_h/ts_l [rows_padded,256] (rows >= B zeroed: the A operand of the modulation GEMM); coef rows (leading dimension
// coef_ld >= 4 floats) receive (c_in, c_skip, c_out, t_hat) in their first four entries.  A conditioning row as the
// denoiser consumes it is kCoefWidth = 8 floats: (c_in, c_skip, c_out, t_hat, t_next, eta, -, -); t_next / eta are
// written by the caller of the sampler and only read by the fused Euler update.
constexpr int kCoefWidth = 8;
cudaError_t launch_time_embed(const float* t_hat, const float* freq, const float* w1, const float* b1,
                              const float* w2, const float* b2, float sigma_data, float* tsilu, __half* ts_h,
                              __half* ts_l, int rows_padded, float* coef, int coef_ld, int B, cudaStream_t st);
// mod[B,Nmod] = tsilu[B,256] * wmod[Nmod,256]^T + bmod   (all AdaLN-Zero linears of the model at once)
cudaError_t launch_mod_gemv(const float* tsilu, const float* wmod, const float* bmod, float* mod, int B,
                            int Nmod, cudaStream_t st);
// x[B*S_pad, c] -> planes: LN_noaffine(x) * (1 + scale) + shift, (shift, scale) = mod[b, off .. off+2c)
cudaError_t launch_adaln(const float* x, const float* mod, int mod_stride, int mod_off, __half* xh, __half* xl,
                         int B, int S_pad, int c, float eps, cudaStream_t st);
// The first AdaLN of the atom encoder fused with AF3DiT.precond (ba rows are produced, stored and normalised in one pass) ...
cudaError_t launch_precond_adaln(const float* x_hat, const float* coef, int coef_stride, const float* a, const float* wx,
                                 const float* bx, float* ba, const float* mod, int mod_stride, int mod_off, __half* xh,
                                 __half* xl, int B, int Na, int S_pad, int c_a, float eps, cudaStream_t st);
// ... and the first AdaLN of the atom decoder fused with AF3DiT.upscale's gather-add (ba += up[atom2tok], stored, normalised)
cudaError_t launch_upscale_adaln(float* ba, const float* up, const int* atom2tok, const float* mod, int mod_stride,
                                 int mod_off, __half* xh, __half* xl, int B, int Na, int Sa_pad, int St_pad, int c_a,
                                 float eps, cudaStream_t st);
// x[rows, c] fp32 -> planes
cudaError_t launch_split(const float* x, __half* xh, __half* xl, size_t n, cudaStream_t st);
// ba[B,S_pad,c_a] = W_x (x_hat * c_in) + b_x + a   (rows >= Na zeroed)
// coef rows are (c_in, c_skip, c_out, t_hat); coef_stride = 4 (one row per sample) or 0 (one row shared by all samples)
cudaError_t launch_precond(const float* x_hat, const float* coef, int coef_stride, const float* a, const float* wx,
                           const float* bx, float* ba, int B, int Na, int S_pad, int c_a, cudaStream_t st);
// bs[B,St_pad,c_s] = segment_sum(h[B,Sa_pad,c_s]) / (n + 1e-3) + s   (rows >= Nt zeroed)
cudaError_t launch_segment_mean(const float* h, const int* tok_start, const float* s, float* bs, int B, int Nt,
                                int Sa_pad, int St_pad, int c_s, cudaStream_t st);
// ba[b, i, :] += up[b, atom2tok[i], :]
cudaError_t launch_gather_add(float* ba, const float* up, const int* atom2tok, int B, int Na, int Sa_pad,
                              int St_pad, int c_a, cudaStream_t st);
// x_den = c_skip * x_hat + c_out * W_r LN(ba); x_next != nullptr additionally writes the physics-free Euler update
// x_next = x_hat + eta * (t_next - t_hat) * (x_hat - x_den) / t_hat   (model.py:263-264,278-281) with t_next, eta taken
// from entries 4, 5 of the sample's conditioning row
cudaError_t launch_denoise_out(const float* ba, const float* x_hat, const float* coef, int coef_stride, const float* ln_w,
                               const float* ln_b, const float* wr, float* x_den, int B, int Na, int S_pad,
                               int c_a, float eps, float* x_next, cudaStream_t st);

// ----------------------------------------------------------------------------- coordinates / physics
cudaError_t launch_centre_augment(const float* x, const float* x_exists, const float* u4, const float* trans,
                                  const float* noise, float lambda, float noise_scale, float trans_scale,
                                  float* x_out, int B, int Na, cudaStream_t st);
cudaError_t launch_euler(const float* x_hat, const float* x_den, const float* aligned, const float* w,
                         const float* t_hat, float t_next, float eta, float* x_next, int B, int Na,
                         cudaStream_t st);
cudaError_t launch_template_eps(const float* x_den, const int* lig_idx, const float* ref_dist, float* eps,
                                int B, int Na, int n_lig, int C, cudaStream_t st);
cudaError_t launch_template_pick(const float* eps, const float* ref_poses, const int* lig_idx, int64_t* used,
                                 float* batch_ref_pos, int B, int Na, int n_lig, int C, cudaStream_t st);
cudaError_t launch_rigid_align(const float* x_pred, const float* x_exists, const float* x_gt, int gt_batched,
                               const float* w, float* aligned, int B, int Na, cudaStream_t st);

// dist[S,S] (fp64) = pairwise RMSD of poses [S,n,3]  (ranking, redocking.py:391)
cudaError_t launch_pairwise_rmsd(const float* poses, double* dist, int S, int n, cudaStream_t st);

// ----------------------------------------------------------------------------- pair-energy physics backend (physics.cu)
struct PairEnergyParams {
    float clash_k;       // clash penalty weight
    float clash_scale;   // clash when d < clash_scale * sig_ij
    float cutoff2;       // squared nonbonded cutoff
    float softcore;      // soft-core constant of the LJ term
};
// Energy of the row atoms in the field of all atoms, per-row energies e_row[B,n_rows], energy[B] (optional) and
// grad[B,Na,3] (written for row atoms only).  partner/p_r0/p_k [Na,E]: bonded-partner table (-1 = empty slot).
cudaError_t launch_pair_energy_grad(const float* x, const float* exists, const float* sigma, const float* eps,
                                    const int* partner, const float* p_r0, const float* p_k, int E, const int* rows,
                                    const unsigned char* in_rows, int n_rows, float* e_row, float* energy, float* grad,
                                    int B, int Na, const PairEnergyParams& pp, cudaStream_t st);
// `iters` descent steps on the row atoms in one launch (rows / in_rows required); cudaErrorInvalidValue when the sample does
// not fit one CTA's shared memory
cudaError_t launch_pair_descend(const float* x, const float* exists, const float* sigma, const float* eps, const int* partner,
                                const float* p_r0, const float* p_k, int E, const int* rows, const unsigned char* in_rows,
                                int n_rows, int iters, float step, float gmax, float* x_out, int B, int Na,
                                const PairEnergyParams& pp, cudaStream_t st);
cudaError_t launch_descent_update(const float* x, const float* grad, const unsigned char* in_rows, float step, float gmax,
                                  float* x_out, int B, int Na, cudaStream_t st);

}  // namespace pdk

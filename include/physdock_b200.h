/* physdock_b200 -- C ABI of the B200-native PhysDock sampling-step library (libphysdock_b200.so).
 *
 * The reference (KexinZhangResearch/PhysDock) has no FFI/plugin interface: its seam is Python-attribute
 * level.  This header is therefore the NEW boundary a maintainer binds with ctypes
 * (see INTEGRATION.md); each entry point names the reference code it replaces (paths relative to the
 * reference repository root).
 *
 * Conventions
 *   - plain C types only: device pointers, sizes, scalars, a CUDA stream passed as void* (cudaStream_t).
 *   - the caller (PyTorch) owns every buffer; the library allocates no device memory and never syncs.
 *     Every call only enqueues kernels on `stream`, so a whole denoising step is CUDA-graph capturable.
 *   - return value 0 = success; otherwise a non-zero code and pdk_last_error() gives the text
 *     (thread-local).  One pdk_dit handle per (device, stream); a handle is not thread-safe.
 *   - fp32 tensors are `const float*`; "planes" are the split-fp16 operand format: a logical fp32 matrix X
 *     is stored as two fp16 matrices (X_hi = fp16(X), X_lo = fp16(X - X_hi)) of the same shape.
 *   - activations are padded per sample to S_pad = round_up(S, 128) rows; pad rows hold finite junk.
 */
#ifndef PHYSDOCK_B200_H
#define PHYSDOCK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PDK_ABI_VERSION 2

int pdk_abi_version(void);
const char* pdk_last_error(void);

/* ------------------------------------------------------------------------------------------------
 * Denoiser handle: AF3DiT.forward (PhysDock/models/layers/transformers.py:235-262)
 * ---------------------------------------------------------------------------------------------- */
typedef struct pdk_dit pdk_dit;

typedef struct {          /* PhysDock/configs.py:59-88 ("dit" section) */
    int64_t c_a, c_ap, c_s, c_z;
    int64_t n_atom_blocks;      /* per atom stack (encoder and decoder each) */
    int64_t n_token_blocks;
    int64_t hidden_a, hidden_s; /* SwiGLU widths, feed_forward.py:18-25 */
    int64_t n_mod;              /* total width of the concatenated AdaLN-Zero linears */
    double sigma_data, eps, inf;
} pdk_dit_dims;

typedef struct {          /* one DiTBlock (transformers.py:149-159); all device pointers */
    const void* wqkv_h; const void* wqkv_l;   /* fp16 [3c, c]: linear_q | linear_k | linear_v rows   (attentions.py:234-236) */
    const void* wo_h;   const void* wo_l;     /* fp16 [c, c]   linear_o                             (attentions.py:240) */
    const void* w13_h;  const void* w13_l;    /* fp16 [2*hidden, c]: w1/w3 rows interleaved in blocks of 16 (feed_forward.py:26-28) */
    const void* w2_h;   const void* w2_l;     /* fp16 [c, hidden] */
    const float* bo;                          /* [c] */
    const float* norm_q; const float* norm_k; /* [32] RMSNorm gains (attentions.py:238-239) */
    int64_t mod_attn_off, mod_ffn_off;        /* column offsets of this block's (shift|scale|gate) in the mod row */
} pdk_block_weights;

typedef struct {
    const float* freq;                        /* [128] exp(-ln(1e4) k/128)  (timestep_embeddings.py:62-67) */
    const float* te_w1; const float* te_b1; const float* te_w2; const float* te_b2;   /* [256,256],[256] x2 */
    const void* wmod_h; const void* wmod_l;   /* fp16 planes [n_mod,256]: every AdaLayerNormZero.linear, concatenated */
    const float* bmod;                        /* [n_mod] */
    const float* wx; const float* bx;         /* linear_x [c_a,3],[c_a] */
    const void* wdown_h; const void* wdown_l; const float* bdown;   /* linear_downscale [c_s,c_a] */
    const void* wup_h;   const void* wup_l;   const float* bup;     /* linear_upscale   [c_a,c_s] */
    const float* norm_r_w; const float* norm_r_b; const float* wr;  /* norm_r, linear_r [3,c_a] */
    const float* wz_atom_T; const float* bz_atom;   /* folded norm_z+linear_z of the 2*n_atom_blocks atom blocks: [c_ap][LH_a], [LH_a] */
    const float* wz_tok_T;  const float* bz_tok;    /* same for the token blocks: [c_z][LH_t], [LH_t] */
    const pdk_block_weights* blocks;          /* host array: encoder blocks, token blocks, decoder blocks */
    int64_t n_blocks;
} pdk_dit_weights;

int pdk_dit_create(const pdk_dit_dims* dims, pdk_dit** out);
int pdk_dit_destroy(pdk_dit* h);
/* copies the pointer table (not the tensors); call again after weights move */
int pdk_dit_set_weights(pdk_dit* h, const pdk_dit_weights* w);

/* padded sequence length used for every activation / bias buffer */
int64_t pdk_pad_len(int64_t n);
/* bytes of the two per-complex pair-bias caches: fp32 [2*n_atom_blocks*H_a, Sa_pad, Sa_pad] and
 * [n_token_blocks*H_s, St_pad, St_pad] */
int pdk_dit_bias_bytes(const pdk_dit* h, int64_t Na, int64_t Nt, size_t* atom_bytes, size_t* token_bytes);
/* bytes of scratch pdk_dit_denoise needs for B samples of the prepared complex */
int pdk_dit_workspace_bytes(const pdk_dit* h, int64_t B, int64_t Na, int64_t Nt, size_t* bytes);

/* Once per complex: caches everything that does not depend on (x_hat, t_hat).
 * Replaces the per-block, per-step `norm_z` + `linear_z` + gen_attn_mask of DiTAttention
 * (attentions.py:246,254-255; utils/tensor_utils.py:642-646) with one pass per pair tensor.
 *   a [Na,c_a], ap [Na,Na,c_ap], s [Nt,c_s], z [Nt,Nt,c_z], ap_mask [Na,Na], z_mask [Nt,Nt]   fp32
 *   tok_start [Nt+1] int32 = exclusive cumsum of batch["token_id_to_chunk_sizes"]
 *   atom2tok  [Na]  int32 = batch["atom_id_to_token_id"]
 * a, s, tok_start, atom2tok and the two bias buffers must stay alive until the next prepare. */
int pdk_dit_prepare_complex(pdk_dit* h, const float* a, const float* ap, const float* s, const float* z,
                            const float* ap_mask, const float* z_mask, const int32_t* tok_start,
                            const int32_t* atom2tok, int64_t Na, int64_t Nt, float* bias_atom,
                            float* bias_tok, void* stream);

/* x_denoised[B,Na,3] = AF3DiT(x_hat[B,Na,3], t_hat[B])  (transformers.py:235-262) for the prepared complex. */
int pdk_dit_denoise(pdk_dit* h, const float* x_hat, const float* t_hat, int64_t B, void* workspace,
                    size_t workspace_bytes, float* x_denoised, void* stream);
/* number of kernels one pdk_dit_denoise call enqueues (for bench.py's gpu_launches) */
int64_t pdk_dit_launches_per_denoise(const pdk_dit* h);

/* Conditioning hoisted out of the step.  Everything AF3DiT derives from t_hat alone -- precond scalars
 * (transformers.py:219-221,229-230), TimestepEmbeddings (timestep_embeddings.py:156-166) and the 36
 * AdaLayerNormZero.linear(SiLU(t)) modulations (adaptive_layer_norm_zero.py:19) -- depends only on the noise level,
 * and the sampler (model.py:211-221) uses ONE noise level per step for all samples, known from the schedule before the
 * loop starts.  pdk_dit_conditioning computes the rows for n noise levels at once (2 launches):
 *   table[i, 0 .. n_mod)            = modulation row of t_hat[i]
 *   table[i, n_mod .. n_mod + 4)    = (c_in, c_skip, c_out, t_hat[i])
 *   table[i, n_mod + 4 .. n_mod + 8) = reserved for the caller: (t_next, eta, -, -), read by the fused Euler update
 * table is fp32 [pad_len(n), table_ld], table_ld >= pdk_dit_cond_width() and a multiple of 4; rows >= n are scratch. */
int64_t pdk_dit_cond_width(const pdk_dit* h);
int pdk_dit_conditioning_workspace_bytes(const pdk_dit* h, int64_t n, size_t* bytes);
int pdk_dit_conditioning(pdk_dit* h, const float* t_hat, int64_t n, void* workspace, size_t workspace_bytes, float* table,
                         int64_t table_ld, void* stream);
/* pdk_dit_denoise with the conditioning rows given: sample b uses cond + b * cond_stride (cond_stride = 0: one row shared
 * by all samples -- the sampler's case).  x_next != NULL additionally writes the physics-free Euler update
 * x_next = x_hat + eta (t_next - t_hat) (x_hat - x_denoised) / t_hat  (model.py:263-264,278-281) from the same kernel
 * that produces x_denoised, with (t_next, eta) taken from the row. */
int pdk_dit_denoise_cond(pdk_dit* h, const float* x_hat, const float* cond, int64_t cond_stride, int64_t B, void* workspace,
                         size_t workspace_bytes, float* x_denoised, float* x_next, void* stream);
int64_t pdk_dit_launches_per_denoise_cond(const pdk_dit* h);

/* ------------------------------------------------------------------------------------------------
 * Sampler-side coordinate / physics ops (PhysDock/models/model.py:211-281)
 * ---------------------------------------------------------------------------------------------- */
/* centre_random_augmentation (utils/tensor_utils.py:576-586) fused with PhysDock.diffuse (model.py:70-85).
 *   u4 [B,4] = the four torch.rand draws (phi0, theta0, phi1, theta1), trans [B,3] = the normal draw,
 *   noise [B,Na,3] or NULL (deterministic steps, t_cur <= gamma_min);
 *   x_out = R (x - masked_mean) + trans_scale*trans + (lambda*noise)*noise_scale,
 *   noise_scale = sqrt(t_hat^2 - t_cur^2) computed by the caller in fp32. */
int pdk_centre_augment(const float* x, const float* x_exists, const float* u4, const float* trans,
                       const float* noise, float lambda, float noise_scale, float trans_scale, float* x_out,
                       int64_t B, int64_t Na, void* stream);
/* d_cur + Euler update (model.py:247-250,263-264,278-281).  aligned/w NULL => d_cur = (x_hat-x_den)/t_hat. */
int pdk_euler_update(const float* x_hat, const float* x_den, const float* aligned, const float* w,
                     const float* t_hat, float t_next, float eta, float* x_next, int64_t B, int64_t Na,
                     void* stream);
/* Template selection (model.py:231-241): eps[B,C] smooth-lDDT mismatch, used[B] = argmin, and
 * batch_ref_pos[b, lig_idx, :] = ref_poses[used[b]].  ref_dist [C,n,n] = pairwise distances of ref_poses [C,n,3]. */
int pdk_template_select(const float* x_den, const int32_t* lig_idx, const float* ref_dist, const float* ref_poses,
                        float* eps, int64_t* used, float* batch_ref_pos, int64_t B, int64_t Na, int64_t n_lig,
                        int64_t C, void* stream);
/* weighted_rigid_align(x_den * x_exists, x_gt, w) (utils/tensor_utils.py:724-778; model.py:245). */
int pdk_rigid_align(const float* x_den, const float* x_exists, const float* x_gt, int gt_batched, const float* w,
                    float* aligned, int64_t B, int64_t Na, void* stream);

/* Host only: the attention kernel's CTA work list for B samples, H heads, n_qtiles 128-row query tiles on n_sms SMs.
 * out[i] = head | q_tile << 8 | first_sample << 16 | n_samples << 24 (n_samples <= 4), ordered by decreasing n_samples;
 * every (head, q tile, sample) appears exactly once.  Returns the number of entries, -1 if the shape needs more than one
 * launch (the library then splits the samples), -2 if cap is too small. */
int64_t pdk_attention_work_list(int64_t B, int64_t H, int64_t n_qtiles, int64_t n_sms, uint32_t* out, int64_t cap);

/* Pose ranking (redocking.py:391): dist[S,S] (fp64) = sqrt(mean over the n ligand atoms of |pose_s - pose_t|^2). */
int pdk_pairwise_rmsd(const float* poses, double* dist, int64_t S, int64_t n, void* stream);

/* Pair-energy physics backend (opt-in, device-resident replacement of get_next_step_pos, model.py:26-52, whose
 * RDKit MMFF94 arithmetic is not part of the reference tree; functional form defined in csrc/physics.cu, oracle =
 * autograd restatement oracle/physdock_oracle.py:pair_energy).
 *   rows [n_rows] / in_rows [Na] (both or neither): the movable atoms (ligand); NULL = every atom, n_rows = Na.
 *   partner / partner_r0 / partner_k [Na,E]: symmetric bonded-partner table, -1 = empty slot; listed pairs are excluded
 *   from the nonbonded terms and carry k (d - r0)^2 when k != 0.
 *   e_row [B,n_rows], energy [B] (may be NULL), grad [B,Na,3] (rows of non-row atoms are not written). */
int pdk_pair_energy_grad(const float* x, const float* x_exists, const float* sigma, const float* eps,
                         const int32_t* partner, const float* partner_r0, const float* partner_k, int64_t E,
                         const int32_t* rows, const uint8_t* in_rows, int64_t n_rows, float clash_k, float clash_scale,
                         float cutoff, float softcore, float* e_row, float* energy, float* grad, int64_t B, int64_t Na,
                         void* stream);
/* `iters` projected-gradient steps x <- x - step * clamp(grad E(x), +-gmax) on the row atoms in ONE launch (one CTA per
 * sample, the sample staged in shared memory; same arithmetic as pdk_pair_energy_grad + pdk_descent_update repeated).
 * Replaces the `maxIters=mmff_iters` minimiser loop of get_next_step_pos (model.py:43).  rows / in_rows are required. */
int pdk_pair_descend(const float* x, const float* x_exists, const float* sigma, const float* eps, const int32_t* partner,
                     const float* partner_r0, const float* partner_k, int64_t E, const int32_t* rows, const uint8_t* in_rows,
                     int64_t n_rows, float clash_k, float clash_scale, float cutoff, float softcore, int64_t iters, float step,
                     float gmax, float* x_out, int64_t B, int64_t Na, void* stream);
/* One projected-gradient step: x_out = x - step * clamp(grad, +-gmax) on the row atoms, copy elsewhere. */
int pdk_descent_update(const float* x, const float* grad, const uint8_t* in_rows, float step, float gmax, float* x_out,
                       int64_t B, int64_t Na, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Op-level entry points (one kernel each) -- what pdk_dit_denoise is built from; exported so the parity
 * tests can check every kernel against the oracle in isolation.
 * ---------------------------------------------------------------------------------------------- */
int pdk_op_pair_bias(const float* pair, const float* mask, const float* wfoldT, const float* bfold, float* bias,
                     int64_t S, int64_t S_pad, int64_t C, int64_t LH, float ln_eps, float inf_, void* stream);
int pdk_op_time_embed(const float* t_hat, const float* freq, const float* w1, const float* b1, const float* w2,
                      const float* b2, float sigma_data, float* tsilu, float* coef, int64_t B, void* stream);
int pdk_op_mod_gemv(const float* tsilu, const float* wmod, const float* bmod, float* mod, int64_t B,
                    int64_t n_mod, void* stream);
int pdk_op_adaln(const float* x, const float* mod, int64_t mod_stride, int64_t mod_off, void* xh, void* xl,
                 int64_t B, int64_t S_pad, int64_t c, float eps, void* stream);
/* the first AdaLN of the atom encoder / decoder also produces its input row: precond (transformers.py:222) resp. the
 * gather-add of upscale (transformers.py:214-216); same arithmetic as pdk_op_precond / pdk_op_gather_add + pdk_op_adaln */
int pdk_op_precond_adaln(const float* x_hat, const float* coef, const float* a, const float* wx, const float* bx, float* ba,
                         const float* mod, int64_t mod_stride, int64_t mod_off, void* xh, void* xl, int64_t B, int64_t Na,
                         int64_t S_pad, int64_t c_a, float eps, void* stream);
int pdk_op_upscale_adaln(float* ba, const float* up, const int32_t* atom2tok, const float* mod, int64_t mod_stride,
                         int64_t mod_off, void* xh, void* xl, int64_t B, int64_t Na, int64_t Sa_pad, int64_t St_pad,
                         int64_t c_a, float eps, void* stream);
int pdk_op_split(const float* x, void* xh, void* xl, int64_t n, void* stream);
int pdk_op_gemm_store(const void* Ah, const void* Al, int64_t lda, const void* Wh, const void* Wl, int64_t ldw,
                      int64_t M, int64_t N, int64_t K, const float* bias, int act_silu, float* out, int64_t ldo,
                      void* stream);
int pdk_op_gemm_gate_resid(const void* Ah, const void* Al, int64_t lda, const void* Wh, const void* Wl, int64_t ldw,
                           int64_t M, int64_t N, int64_t K, const float* bias, const float* gate,
                           int64_t gate_stride, int64_t rows_per_sample, float* x, int64_t ldx, void* stream);
int pdk_op_gemm_swiglu(const void* Ah, const void* Al, int64_t lda, const void* Wh, const void* Wl, int64_t ldw,
                       int64_t M, int64_t N, int64_t K, void* ph, void* pl, int64_t ldp, void* stream);
/* q, k, v: fp16 [B,H,S_pad,64], each row = [hi 32 | lo 32] (the two split planes interleaved: 128-byte rows) */
/* Fused DiTTransition of the atom stacks (c = 128; transitions.py:21-30): x[M,128] += w2(SiLU(w1 xn) * w3 xn) * gate,
 * xn = LN(x) * (1 + scale) + shift, (shift | scale | gate) = mod[sample * mod_stride + mod_off ...]. */
int pdk_op_transition_fused(float* x, const float* mod, int64_t mod_stride, int64_t mod_off, const void* w13h,
                            const void* w13l, const void* w2h, const void* w2l, int64_t M, int64_t hidden,
                            int64_t rows_per_sample, float eps, void* stream);
int pdk_op_gemm_qkv(const void* Ah, const void* Al, int64_t lda, const void* Wh, const void* Wl, int64_t ldw,
                    int64_t M, int64_t c, const float* norm_q, const float* norm_k, float rms_eps, float q_scale,
                    int64_t rows_per_sample, void* q, void* k, void* v, void* stream);
int pdk_op_attention(const void* q, const void* k, const void* v, const float* bias, void* oh, void* ol, int64_t B,
                     int64_t H, int64_t S_pad, void* stream);
int pdk_op_precond(const float* x_hat, const float* coef, const float* a, const float* wx, const float* bx, float* ba,
                   int64_t B, int64_t Na, int64_t S_pad, int64_t c_a, void* stream);
int pdk_op_segment_mean(const float* h, const int32_t* tok_start, const float* s, float* bs, int64_t B, int64_t Nt,
                        int64_t Sa_pad, int64_t St_pad, int64_t c_s, void* stream);
int pdk_op_gather_add(float* ba, const float* up, const int32_t* atom2tok, int64_t B, int64_t Na, int64_t Sa_pad,
                      int64_t St_pad, int64_t c_a, void* stream);
int pdk_op_denoise_out(const float* ba, const float* x_hat, const float* coef, const float* ln_w, const float* ln_b,
                       const float* wr, float* x_den, int64_t B, int64_t Na, int64_t S_pad, int64_t c_a, float eps,
                       void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PHYSDOCK_B200_H */

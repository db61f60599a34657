// C ABI (include/physdock_b200.h): handle, workspace carve-up, the AF3DiT pipeline, op-level wrappers.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/physdock_b200.h"
#include "common.cuh"
#include "kernels.h"

using namespace pdk;

namespace {

thread_local std::string g_err;

int fail(const char* where, cudaError_t e) {
    g_err = std::string(where) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
    return (int)e ? (int)e : -1;
}
int fail_msg(const char* where, const char* msg) {
    g_err = std::string(where) + ": " + msg;
    return -1;
}
#ifdef PDK_DBG_SKIP
// Debug builds only (tools/gemm_variants.sh SKIP): PDK_SKIP="adaln,time_embed,..." drops the launches whose label contains one
// of the words, to measure the upper bound of fusing / removing them (results are then wrong by construction).
inline bool dbg_skip(const char* where) {
    static const char* list = getenv("PDK_SKIP");
    if (!list) return false;
    std::string l(list), w(where);
    size_t a = 0;
    while (a <= l.size()) {
        size_t b = l.find(',', a);
        if (b == std::string::npos) b = l.size();
        if (b > a && w.find(l.substr(a, b - a)) != std::string::npos) return true;
        a = b + 1;
    }
    return false;
}
#define PDK_TRY(where, expr)                          \
    do {                                              \
        if (dbg_skip(where)) break;                   \
        cudaError_t _e = (expr);                      \
        if (_e != cudaSuccess) return fail(where, _e); \
    } while (0)
#else
#define PDK_TRY(where, expr)                          \
    do {                                              \
        cudaError_t _e = (expr);                      \
        if (_e != cudaSuccess) return fail(where, _e); \
    } while (0)
#endif

inline int64_t pad128(int64_t n) { return (n + 127) / 128 * 128; }
inline size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }
inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }
inline const __half* H(const void* p) { return reinterpret_cast<const __half*>(p); }
inline __half* H(void* p) { return reinterpret_cast<__half*>(p); }

}  // namespace

struct pdk_dit {
    pdk_dit_dims d{};
    pdk_dit_weights w{};
    std::vector<pdk_block_weights> blocks;
    bool have_weights = false;
    // prepared complex
    bool prepared = false;
    const float* a = nullptr;
    const float* s = nullptr;
    const int32_t* tok_start = nullptr;
    const int32_t* atom2tok = nullptr;
    const float* bias_atom = nullptr;
    const float* bias_tok = nullptr;
    int64_t Na = 0, Nt = 0, Sa = 0, St = 0;   // Sa/St = padded lengths
    int H_a() const { return (int)(d.c_a / kHeadDim); }
    int H_s() const { return (int)(d.c_s / kHeadDim); }
};

namespace {

struct Workspace {
    float *coef, *tsilu, *mod, *ba, *bs, *down, *up;
    __half *xh, *xl, *q, *k, *v, *oh, *ol, *hh, *hl, *tsh, *tsl;
    size_t bytes;
};

// Carves the scratch buffer.  base == nullptr just measures.
Workspace carve(const pdk_dit& h, int64_t B, int64_t Sa, int64_t St, uint8_t* base) {
    Workspace w{};
    size_t off = 0;
    auto take = [&](size_t bytes) {
        uint8_t* p = base ? base + off : nullptr;
        off += align_up(bytes);
        return p;
    };
    const size_t Ma = (size_t)B * Sa, Mt = (size_t)B * St;
    const size_t act = std::max(Ma * h.d.c_a, Mt * h.d.c_s);            // elements of the widest activation
    const size_t hid = std::max(Ma * h.d.hidden_a, Mt * h.d.hidden_s);
    w.coef = (float*)take((size_t)B * kCoefWidth * 4);
    const size_t Bp = (size_t)pad128(B);                    // the modulation GEMM runs on 128-row tiles
    w.tsilu = (float*)take((size_t)B * kTimeDim * 4);
    w.tsh = (__half*)take(Bp * kTimeDim * 2);
    w.tsl = (__half*)take(Bp * kTimeDim * 2);
    w.mod = (float*)take(Bp * h.d.n_mod * 4);
    w.ba = (float*)take(Ma * h.d.c_a * 4);
    w.bs = (float*)take(Mt * h.d.c_s * 4);
    w.down = (float*)take(Ma * h.d.c_s * 4);
    w.up = (float*)take(Mt * h.d.c_a * 4);
    __half** planes[] = {&w.xh, &w.xl, &w.oh, &w.ol};
    for (auto pp : planes) *pp = (__half*)take(act * 2);
    __half** qkv[] = {&w.q, &w.k, &w.v};                 // interleaved [hi|lo]: twice the halves
    for (auto pp : qkv) *pp = (__half*)take(act * 4);
    w.hh = (__half*)take(hid * 2);
    w.hl = (__half*)take(hid * 2);
    w.bytes = off;
    return w;
}

constexpr int kLaunchesPerBlock = 7;
constexpr int kLaunchesPerFusedBlock = 5;        // atom blocks: the transition is one kernel (transition_umma.cu)
inline bool fused_transition_enabled() {
    static const bool on = !measure_switch("PDK_NO_FUSED_TRANSITION");
    return on;
}

// One DiTBlock (transformers.py:155-159): x += Attn(x); x += Transition(x).
// `mod` = modulation rows (all AdaLN-Zero layers side by side), `nmod` = stride between the rows of consecutive samples
// in floats (0: one row shared by every sample, the sampler's case: all samples of a step have the same noise level).
// first_adaln_done: the caller already produced x and its attention-AdaLN planes (launch_precond_adaln / launch_upscale_adaln).
int run_block(const pdk_dit& h, const pdk_block_weights& bw, const Workspace& ws, const float* mod, int nmod, float* x,
              int64_t B, int64_t Sp, int c, int hidden, const float* bias, cudaStream_t st, bool first_adaln_done = false) {
    const int M = (int)(B * Sp);
    const int Hh = c / kHeadDim;
    const float eps = (float)h.d.eps;
    // --- attention (attentions.py:240-265)
    if (!first_adaln_done)
        PDK_TRY("adaln(attn)", launch_adaln(x, mod, nmod, (int)bw.mod_attn_off, ws.xh, ws.xl, (int)B, (int)Sp, c, eps, st));
    GemmArgs g{};
    g.Ah = ws.xh; g.Al = ws.xl; g.lda = c;
    g.Wh = H(bw.wqkv_h); g.Wl = H(bw.wqkv_l); g.ldw = c;
    g.M = M; g.N = 3 * c; g.K = c;
    g.q = ws.q; g.k = ws.k; g.v = ws.v;
    g.norm_q = bw.norm_q; g.norm_k = bw.norm_k; g.c = c; g.rows_per_sample = (int)Sp;
    g.rms_eps = eps; g.q_scale = kLog2e / sqrtf((float)kHeadDim);
    PDK_TRY("gemm(qkv)", launch_gemm(EPI_QKV, g, st));
    AttnArgs at{ws.q, ws.k, ws.v, bias, ws.oh, ws.ol, (int)B, Hh, (int)Sp, c, 0, nullptr};
    PDK_TRY("attention", launch_attention(at, st));
    g = GemmArgs{};
    g.Ah = ws.oh; g.Al = ws.ol; g.lda = c;
    g.Wh = H(bw.wo_h); g.Wl = H(bw.wo_l); g.ldw = c;
    g.M = M; g.N = c; g.K = c;
    g.bias = bw.bo; g.out = x; g.ldo = c;
    g.gate = mod + bw.mod_attn_off + 2 * c; g.gate_stride = nmod; g.rows_per_sample = (int)Sp;
    PDK_TRY("gemm(out)", launch_gemm(EPI_GATE_RESID, g, st));
    // --- transition (transitions.py:21-30)
    if (c == 128 && fused_transition_enabled()) {
        TransitionArgs ta{};
        ta.x = x; ta.mod = mod; ta.mod_stride = nmod; ta.mod_off = (int)bw.mod_ffn_off;
        ta.w13h = H(bw.w13_h); ta.w13l = H(bw.w13_l); ta.w2h = H(bw.w2_h); ta.w2l = H(bw.w2_l);
        ta.M = M; ta.hidden = hidden; ta.rows_per_sample = (int)Sp; ta.eps = eps;
        PDK_TRY("transition(fused)", launch_transition_fused(ta, st));
        return 0;
    }
    PDK_TRY("adaln(ffn)", launch_adaln(x, mod, nmod, (int)bw.mod_ffn_off, ws.xh, ws.xl, (int)B, (int)Sp, c, eps, st));
    g = GemmArgs{};
    g.Ah = ws.xh; g.Al = ws.xl; g.lda = c;
    g.Wh = H(bw.w13_h); g.Wl = H(bw.w13_l); g.ldw = c;
    g.M = M; g.N = 2 * hidden; g.K = c;
    g.ph = ws.hh; g.pl = ws.hl; g.ldp = hidden;
    PDK_TRY("gemm(w13)", launch_gemm(EPI_SWIGLU, g, st));
    g = GemmArgs{};
    g.Ah = ws.hh; g.Al = ws.hl; g.lda = hidden;
    g.Wh = H(bw.w2_h); g.Wl = H(bw.w2_l); g.ldw = hidden;
    g.M = M; g.N = c; g.K = hidden;
    g.out = x; g.ldo = c;
    g.gate = mod + bw.mod_ffn_off + 2 * c; g.gate_stride = nmod; g.rows_per_sample = (int)Sp;
    PDK_TRY("gemm(w2)", launch_gemm(EPI_GATE_RESID, g, st));
    return 0;
}

}  // namespace

extern "C" {

int pdk_abi_version(void) { return PDK_ABI_VERSION; }
#ifdef PDK_MEASURE
// measurement builds only (not in the public header): per-unit timeline of the attention kernel, tools/trace_attention.py
void pdk_debug_attention_trace(void* buf) { set_attention_trace(reinterpret_cast<long long*>(buf)); }
#endif
const char* pdk_last_error(void) { return g_err.c_str(); }
int64_t pdk_pad_len(int64_t n) { return pad128(n); }

int pdk_dit_create(const pdk_dit_dims* dims, pdk_dit** out) {
    if (!dims || !out) return fail_msg("pdk_dit_create", "null argument");
    const pdk_dit_dims& d = *dims;
    if (d.c_a != 128 || d.c_s != 512 || d.c_ap != 16 || d.c_z != 128)
        return fail_msg("pdk_dit_create", "kernels are specialised for c_a=128, c_ap=16, c_s=512, c_z=128 (PhysDock/configs.py:59-63)");
    if (d.hidden_a % 32 || d.hidden_s % 32 || d.hidden_a % 64 || d.hidden_s % 64)
        return fail_msg("pdk_dit_create", "SwiGLU widths must be multiples of 64");
    if (d.n_atom_blocks <= 0 || d.n_token_blocks <= 0 || d.n_mod <= 0 || d.n_mod % 128)
        return fail_msg("pdk_dit_create", "bad block counts / n_mod");
    pdk_dit* h = new (std::nothrow) pdk_dit();
    if (!h) return fail_msg("pdk_dit_create", "out of host memory");
    h->d = d;
    *out = h;
    return 0;
}

int pdk_dit_destroy(pdk_dit* h) {
    delete h;
    return 0;
}

int pdk_dit_set_weights(pdk_dit* h, const pdk_dit_weights* w) {
    if (!h || !w || !w->blocks) return fail_msg("pdk_dit_set_weights", "null argument");
    const int64_t want = 2 * h->d.n_atom_blocks + h->d.n_token_blocks;
    if (w->n_blocks != want) return fail_msg("pdk_dit_set_weights", "n_blocks != 2*n_atom_blocks + n_token_blocks");
    h->w = *w;
    h->blocks.assign(w->blocks, w->blocks + w->n_blocks);
    h->w.blocks = h->blocks.data();
    h->have_weights = true;
    return 0;
}

int pdk_dit_bias_bytes(const pdk_dit* h, int64_t Na, int64_t Nt, size_t* atom_bytes, size_t* token_bytes) {
    if (!h || !atom_bytes || !token_bytes || Na <= 0 || Nt <= 0) return fail_msg("pdk_dit_bias_bytes", "bad argument");
    const size_t Sa = pad128(Na), St = pad128(Nt);
    *atom_bytes = (size_t)(2 * h->d.n_atom_blocks) * h->H_a() * Sa * Sa * 4;
    *token_bytes = (size_t)h->d.n_token_blocks * h->H_s() * St * St * 4;
    return 0;
}

int pdk_dit_workspace_bytes(const pdk_dit* h, int64_t B, int64_t Na, int64_t Nt, size_t* bytes) {
    if (!h || !bytes || B <= 0 || Na <= 0 || Nt <= 0) return fail_msg("pdk_dit_workspace_bytes", "bad argument");
    *bytes = carve(*h, B, pad128(Na), pad128(Nt), nullptr).bytes;
    return 0;
}

int pdk_dit_prepare_complex(pdk_dit* h, const float* a, const float* ap, const float* s, const float* z,
                            const float* ap_mask, const float* z_mask, const int32_t* tok_start,
                            const int32_t* atom2tok, int64_t Na, int64_t Nt, float* bias_atom,
                            float* bias_tok, void* stream) {
    if (!h || !h->have_weights) return fail_msg("pdk_dit_prepare_complex", "weights not set");
    if (!a || !ap || !s || !z || !ap_mask || !z_mask || !tok_start || !atom2tok || !bias_atom || !bias_tok)
        return fail_msg("pdk_dit_prepare_complex", "null argument");
    if (Na <= 0 || Nt <= 0) return fail_msg("pdk_dit_prepare_complex", "empty complex");
    h->prepared = false;
    const int64_t Sa = pad128(Na), St = pad128(Nt);
    const int LHa = (int)(2 * h->d.n_atom_blocks) * h->H_a();
    const int LHt = (int)h->d.n_token_blocks * h->H_s();
    // nn.LayerNorm default eps (attentions.py:232 passes none)
    PDK_TRY("pair_bias(atom)", launch_pair_bias(ap, ap_mask, h->w.wz_atom_T, h->w.bz_atom, bias_atom, (int)Na, (int)Sa,
                                                (int)h->d.c_ap, LHa, 1e-5f, (float)h->d.inf, S(stream)));
    PDK_TRY("pair_bias(token)", launch_pair_bias(z, z_mask, h->w.wz_tok_T, h->w.bz_tok, bias_tok, (int)Nt, (int)St,
                                                 (int)h->d.c_z, LHt, 1e-5f, (float)h->d.inf, S(stream)));
    h->a = a; h->s = s; h->tok_start = tok_start; h->atom2tok = atom2tok;
    h->bias_atom = bias_atom; h->bias_tok = bias_tok;
    h->Na = Na; h->Nt = Nt; h->Sa = Sa; h->St = St;
    h->prepared = true;
    return 0;
}

int64_t pdk_dit_launches_per_denoise(const pdk_dit* h) {
    if (!h) return 0;
    // time_embed, mod | blocks (precond and the upscale gather-add ride in the first AdaLN of their atom stack) |
    // split, gemm(down), segmean | split, gemm(up) | denoise_out
    const int64_t per_atom_block = fused_transition_enabled() ? kLaunchesPerFusedBlock : kLaunchesPerBlock;
    return 2 + per_atom_block * 2 * h->d.n_atom_blocks + kLaunchesPerBlock * h->d.n_token_blocks + 3 + 2 + 1;
}
// pdk_dit_denoise_cond: the two conditioning launches are not part of the step
int64_t pdk_dit_launches_per_denoise_cond(const pdk_dit* h) { return h ? pdk_dit_launches_per_denoise(h) - 2 : 0; }

// time embedding + all AdaLN-Zero modulations for n noise levels: table rows [mod (n_mod) | coef (kCoefWidth)]
static int run_conditioning(pdk_dit* h, const float* t_hat, int64_t n, __half* tsh, __half* tsl, float* table, int64_t table_ld,
                            cudaStream_t st) {
    const pdk_dit_dims& d = h->d;
    // precond scalars + TimestepEmbeddings (transformers.py:218-226) ...
    PDK_TRY("time_embed", launch_time_embed(t_hat, h->w.freq, h->w.te_w1, h->w.te_b1, h->w.te_w2, h->w.te_b2, (float)d.sigma_data,
                                            nullptr, tsh, tsl, (int)pad128(n), table + d.n_mod, (int)table_ld, (int)n, st));
    // ... and every AdaLayerNormZero.linear of the model as ONE tensor-core GEMM [pad128(n) x 256] x [n_mod x 256]^T
    // (adaptive_layer_norm_zero.py:19, all 36 layers)
    GemmArgs g{};
    g.Ah = tsh; g.Al = tsl; g.lda = kTimeDim;
    g.Wh = H(h->w.wmod_h); g.Wl = H(h->w.wmod_l); g.ldw = kTimeDim;
    g.M = (int)pad128(n); g.N = (int)d.n_mod; g.K = kTimeDim;
    g.bias = h->w.bmod; g.out = table; g.ldo = (int)table_ld;
    PDK_TRY("gemm(mod)", launch_gemm(EPI_STORE, g, st));
    return 0;
}

// AF3DiT.forward (transformers.py:235-262) given the conditioning rows.
static int run_denoise(pdk_dit* h, const float* x_hat, const float* mod, int mod_stride, const float* coef, int coef_stride,
                       int64_t B, const Workspace& ws, float* x_denoised, float* x_next, cudaStream_t st) {
    const pdk_dit_dims& d = h->d;
    const int64_t Sa = h->Sa, St = h->St;
    const int ca = (int)d.c_a, cs = (int)d.c_s;
    const int Ha = h->H_a(), Hs = h->H_s();
    const int nA = (int)d.n_atom_blocks, nT = (int)d.n_token_blocks;
    // precond (transformers.py:218-226) fused with the first AdaLN of the atom encoder: ba is produced, stored and normalised
    // in one pass
    PDK_TRY("precond+adaln", launch_precond_adaln(x_hat, coef, coef_stride, h->a, h->w.wx, h->w.bx, ws.ba, mod, mod_stride,
                                                  (int)h->blocks[0].mod_attn_off, ws.xh, ws.xl, (int)B, (int)h->Na, (int)Sa, ca,
                                                  (float)d.eps, st));

    // atom encoder (transformers.py:252)
    const size_t plane_a = (size_t)Ha * Sa * Sa, plane_t = (size_t)Hs * St * St;
    for (int l = 0; l < nA; ++l) {
        int rc = run_block(*h, h->blocks[l], ws, mod, mod_stride, ws.ba, B, Sa, ca, (int)d.hidden_a, h->bias_atom + l * plane_a, st,
                           l == 0);
        if (rc) return rc;
    }
    // downscale (transformers.py:205-212)
    PDK_TRY("split(ba)", launch_split(ws.ba, ws.xh, ws.xl, (size_t)B * Sa * ca, st));
    {
        GemmArgs g{};
        g.Ah = ws.xh; g.Al = ws.xl; g.lda = ca;
        g.Wh = H(h->w.wdown_h); g.Wl = H(h->w.wdown_l); g.ldw = ca;
        g.M = (int)(B * Sa); g.N = cs; g.K = ca;
        g.bias = h->w.bdown; g.act_silu = 1; g.out = ws.down; g.ldo = cs;
        PDK_TRY("gemm(down)", launch_gemm(EPI_STORE, g, st));
    }
    PDK_TRY("segment_mean", launch_segment_mean(ws.down, h->tok_start, h->s, ws.bs, (int)B, (int)h->Nt, (int)Sa, (int)St, cs, st));
    // token DiT (transformers.py:255)
    for (int l = 0; l < nT; ++l) {
        int rc = run_block(*h, h->blocks[nA + l], ws, mod, mod_stride, ws.bs, B, St, cs, (int)d.hidden_s, h->bias_tok + l * plane_t, st);
        if (rc) return rc;
    }
    // upscale (transformers.py:214-216)
    PDK_TRY("split(bs)", launch_split(ws.bs, ws.xh, ws.xl, (size_t)B * St * cs, st));
    {
        GemmArgs g{};
        g.Ah = ws.xh; g.Al = ws.xl; g.lda = cs;
        g.Wh = H(h->w.wup_h); g.Wl = H(h->w.wup_l); g.ldw = cs;
        g.M = (int)(B * St); g.N = ca; g.K = cs;
        g.bias = h->w.bup; g.out = ws.up; g.ldo = ca;
        PDK_TRY("gemm(up)", launch_gemm(EPI_STORE, g, st));
    }
    // ... the gather-add ba += up[atom_id_to_token_id] runs inside the first AdaLN of the atom decoder
    PDK_TRY("upscale+adaln", launch_upscale_adaln(ws.ba, ws.up, h->atom2tok, mod, mod_stride, (int)h->blocks[nA + nT].mod_attn_off,
                                                  ws.xh, ws.xl, (int)B, (int)h->Na, (int)Sa, (int)St, ca, (float)d.eps, st));
    // atom decoder (transformers.py:259)
    for (int l = 0; l < nA; ++l) {
        int rc = run_block(*h, h->blocks[nA + nT + l], ws, mod, mod_stride, ws.ba, B, Sa, ca, (int)d.hidden_a,
                           h->bias_atom + (size_t)(nA + l) * plane_a, st, l == 0);
        if (rc) return rc;
    }
    // denoise (transformers.py:228-233) [+ the physics-free Euler update, model.py:263-264,278-281]
    PDK_TRY("denoise_out", launch_denoise_out(ws.ba, x_hat, coef, coef_stride, h->w.norm_r_w, h->w.norm_r_b, h->w.wr, x_denoised,
                                              (int)B, (int)h->Na, (int)Sa, ca, (float)d.eps, x_next, st));
    return 0;
}

int pdk_dit_denoise(pdk_dit* h, const float* x_hat, const float* t_hat, int64_t B, void* workspace,
                    size_t workspace_bytes, float* x_denoised, void* stream) {
    if (!h || !h->prepared) return fail_msg("pdk_dit_denoise", "no prepared complex");
    if (!x_hat || !t_hat || !workspace || !x_denoised || B <= 0) return fail_msg("pdk_dit_denoise", "bad argument");
    Workspace ws = carve(*h, B, h->Sa, h->St, reinterpret_cast<uint8_t*>(workspace));
    if (ws.bytes > workspace_bytes) return fail_msg("pdk_dit_denoise", "workspace too small");
    // per-sample noise levels: conditioning rows computed here (2 launches), mod [pad128(B), n_mod], coef [B, kCoefWidth]
    cudaStream_t st = S(stream);
    const pdk_dit_dims& d = h->d;
    PDK_TRY("time_embed", launch_time_embed(t_hat, h->w.freq, h->w.te_w1, h->w.te_b1, h->w.te_w2, h->w.te_b2, (float)d.sigma_data,
                                            nullptr, ws.tsh, ws.tsl, (int)pad128(B), ws.coef, kCoefWidth, (int)B, st));
    {
        GemmArgs g{};
        g.Ah = ws.tsh; g.Al = ws.tsl; g.lda = kTimeDim;
        g.Wh = H(h->w.wmod_h); g.Wl = H(h->w.wmod_l); g.ldw = kTimeDim;
        g.M = (int)pad128(B); g.N = (int)d.n_mod; g.K = kTimeDim;
        g.bias = h->w.bmod; g.out = ws.mod; g.ldo = (int)d.n_mod;
        PDK_TRY("gemm(mod)", launch_gemm(EPI_STORE, g, st));
    }
    return run_denoise(h, x_hat, ws.mod, (int)d.n_mod, ws.coef, kCoefWidth, B, ws, x_denoised, nullptr, st);
}

int64_t pdk_dit_cond_width(const pdk_dit* h) { return h ? h->d.n_mod + kCoefWidth : 0; }

int pdk_dit_conditioning_workspace_bytes(const pdk_dit* h, int64_t n, size_t* bytes) {
    if (!h || !bytes || n <= 0) return fail_msg("pdk_dit_conditioning_workspace_bytes", "bad argument");
    *bytes = 2 * align_up((size_t)pad128(n) * kTimeDim * 2);
    return 0;
}

int pdk_dit_conditioning(pdk_dit* h, const float* t_hat, int64_t n, void* workspace, size_t workspace_bytes, float* table,
                         int64_t table_ld, void* stream) {
    if (!h || !h->have_weights) return fail_msg("pdk_dit_conditioning", "weights not set");
    if (!t_hat || !workspace || !table || n <= 0) return fail_msg("pdk_dit_conditioning", "bad argument");
    if (table_ld < h->d.n_mod + kCoefWidth || table_ld % 4) return fail_msg("pdk_dit_conditioning", "table_ld < n_mod + 8 or not a multiple of 4");
    const size_t plane = align_up((size_t)pad128(n) * kTimeDim * 2);
    if (2 * plane > workspace_bytes) return fail_msg("pdk_dit_conditioning", "workspace too small");
    uint8_t* base = reinterpret_cast<uint8_t*>(workspace);
    return run_conditioning(h, t_hat, n, reinterpret_cast<__half*>(base), reinterpret_cast<__half*>(base + plane), table, table_ld,
                            S(stream));
}

int pdk_dit_denoise_cond(pdk_dit* h, const float* x_hat, const float* cond, int64_t cond_stride, int64_t B, void* workspace,
                         size_t workspace_bytes, float* x_denoised, float* x_next, void* stream) {
    if (!h || !h->prepared) return fail_msg("pdk_dit_denoise_cond", "no prepared complex");
    if (!x_hat || !cond || !workspace || !x_denoised || B <= 0) return fail_msg("pdk_dit_denoise_cond", "bad argument");
    if (cond_stride != 0 && (cond_stride < h->d.n_mod + kCoefWidth || cond_stride % 4))
        return fail_msg("pdk_dit_denoise_cond", "cond_stride must be 0 (shared row) or >= n_mod + 8 and a multiple of 4");
    Workspace ws = carve(*h, B, h->Sa, h->St, reinterpret_cast<uint8_t*>(workspace));
    if (ws.bytes > workspace_bytes) return fail_msg("pdk_dit_denoise_cond", "workspace too small");
    return run_denoise(h, x_hat, cond, (int)cond_stride, cond + h->d.n_mod, (int)cond_stride, B, ws, x_denoised, x_next, S(stream));
}

// ---------------------------------------------------------------------------------- sampler ops
int pdk_centre_augment(const float* x, const float* x_exists, const float* u4, const float* trans,
                       const float* noise, float lambda, float noise_scale, float trans_scale, float* x_out,
                       int64_t B, int64_t Na, void* stream) {
    if (!x || !x_exists || !u4 || !trans || !x_out) return fail_msg("pdk_centre_augment", "null argument");
    PDK_TRY("centre_augment", launch_centre_augment(x, x_exists, u4, trans, noise, lambda, noise_scale, trans_scale,
                                                    x_out, (int)B, (int)Na, S(stream)));
    return 0;
}

int pdk_euler_update(const float* x_hat, const float* x_den, const float* aligned, const float* w,
                     const float* t_hat, float t_next, float eta, float* x_next, int64_t B, int64_t Na,
                     void* stream) {
    if (!x_hat || !x_den || !t_hat || !x_next) return fail_msg("pdk_euler_update", "null argument");
    PDK_TRY("euler", launch_euler(x_hat, x_den, aligned, w, t_hat, t_next, eta, x_next, (int)B, (int)Na, S(stream)));
    return 0;
}

int pdk_template_select(const float* x_den, const int32_t* lig_idx, const float* ref_dist, const float* ref_poses,
                        float* eps, int64_t* used, float* batch_ref_pos, int64_t B, int64_t Na, int64_t n_lig,
                        int64_t C, void* stream) {
    if (!x_den || !lig_idx || !ref_dist || !ref_poses || !eps || !used || !batch_ref_pos)
        return fail_msg("pdk_template_select", "null argument");
    PDK_TRY("template_eps", launch_template_eps(x_den, lig_idx, ref_dist, eps, (int)B, (int)Na, (int)n_lig, (int)C, S(stream)));
    PDK_TRY("template_pick", launch_template_pick(eps, ref_poses, lig_idx, used, batch_ref_pos, (int)B, (int)Na, (int)n_lig,
                                                  (int)C, S(stream)));
    return 0;
}

int64_t pdk_attention_work_list(int64_t B, int64_t H, int64_t n_qtiles, int64_t n_sms, uint32_t* out, int64_t cap) {
    if (!out || B <= 0 || H <= 0 || n_qtiles <= 0 || n_sms <= 0) { fail_msg("pdk_attention_work_list", "bad argument"); return -3; }
    return attention_work_list((int)B, (int)H, (int)n_qtiles, (int)n_sms, out, (int)cap);
}

int pdk_pairwise_rmsd(const float* poses, double* dist, int64_t n_poses, int64_t n, void* stream) {
    if (!poses || !dist) return fail_msg("pdk_pairwise_rmsd", "null argument");
    PDK_TRY("pairwise_rmsd", launch_pairwise_rmsd(poses, dist, (int)n_poses, (int)n, S(stream)));
    return 0;
}

int pdk_pair_energy_grad(const float* x, const float* x_exists, const float* sigma, const float* eps,
                         const int32_t* partner, const float* partner_r0, const float* partner_k, int64_t E,
                         const int32_t* rows, const uint8_t* in_rows, int64_t n_rows, float clash_k, float clash_scale,
                         float cutoff, float softcore, float* e_row, float* energy, float* grad, int64_t B, int64_t Na,
                         void* stream) {
    if (!x || !x_exists || !sigma || !eps || !e_row || !grad) return fail_msg("pdk_pair_energy_grad", "null argument");
    if (E > 0 && (!partner || !partner_r0 || !partner_k)) return fail_msg("pdk_pair_energy_grad", "partner table missing");
    if ((rows == nullptr) != (in_rows == nullptr)) return fail_msg("pdk_pair_energy_grad", "rows and in_rows go together");
    if (rows == nullptr && n_rows != Na) return fail_msg("pdk_pair_energy_grad", "n_rows must equal Na when rows is NULL");
    const PairEnergyParams pp{clash_k, clash_scale, cutoff * cutoff, softcore};
    PDK_TRY("pair_energy_grad", launch_pair_energy_grad(x, x_exists, sigma, eps, partner, partner_r0, partner_k, (int)E, rows,
                                                        in_rows, (int)n_rows, e_row, energy, grad, (int)B, (int)Na, pp, S(stream)));
    return 0;
}

int pdk_pair_descend(const float* x, const float* x_exists, const float* sigma, const float* eps, const int32_t* partner,
                     const float* partner_r0, const float* partner_k, int64_t E, const int32_t* rows, const uint8_t* in_rows,
                     int64_t n_rows, float clash_k, float clash_scale, float cutoff, float softcore, int64_t iters, float step,
                     float gmax, float* x_out, int64_t B, int64_t Na, void* stream) {
    if (!x || !x_exists || !sigma || !eps || !rows || !in_rows || !x_out) return fail_msg("pdk_pair_descend", "null argument");
    if (E > 0 && (!partner || !partner_r0 || !partner_k)) return fail_msg("pdk_pair_descend", "partner table missing");
    const PairEnergyParams pp{clash_k, clash_scale, cutoff * cutoff, softcore};
    PDK_TRY("pair_descend", launch_pair_descend(x, x_exists, sigma, eps, partner, partner_r0, partner_k, (int)E, rows, in_rows,
                                                (int)n_rows, (int)iters, step, gmax, x_out, (int)B, (int)Na, pp, S(stream)));
    return 0;
}

int pdk_descent_update(const float* x, const float* grad, const uint8_t* in_rows, float step, float gmax, float* x_out,
                       int64_t B, int64_t Na, void* stream) {
    if (!x || !grad || !x_out) return fail_msg("pdk_descent_update", "null argument");
    PDK_TRY("descent_update", launch_descent_update(x, grad, in_rows, step, gmax, x_out, (int)B, (int)Na, S(stream)));
    return 0;
}

int pdk_rigid_align(const float* x_den, const float* x_exists, const float* x_gt, int gt_batched, const float* w,
                    float* aligned, int64_t B, int64_t Na, void* stream) {
    if (!x_den || !x_exists || !x_gt || !w || !aligned) return fail_msg("pdk_rigid_align", "null argument");
    PDK_TRY("rigid_align", launch_rigid_align(x_den, x_exists, x_gt, gt_batched, w, aligned, (int)B, (int)Na, S(stream)));
    return 0;
}

// ---------------------------------------------------------------------------------- op-level wrappers
int pdk_op_pair_bias(const float* pair, const float* mask, const float* wfoldT, const float* bfold, float* bias,
                     int64_t Sn, int64_t S_pad, int64_t C, int64_t LH, float ln_eps, float inf_, void* stream) {
    PDK_TRY("pair_bias", launch_pair_bias(pair, mask, wfoldT, bfold, bias, (int)Sn, (int)S_pad, (int)C, (int)LH, ln_eps, inf_, S(stream)));
    return 0;
}
int pdk_op_time_embed(const float* t_hat, const float* freq, const float* w1, const float* b1, const float* w2,
                      const float* b2, float sigma_data, float* tsilu, float* coef, int64_t B, void* stream) {
    PDK_TRY("time_embed", launch_time_embed(t_hat, freq, w1, b1, w2, b2, sigma_data, tsilu, nullptr, nullptr, 0, coef, 4, (int)B, S(stream)));
    return 0;
}
int pdk_op_mod_gemv(const float* tsilu, const float* wmod, const float* bmod, float* mod, int64_t B, int64_t n_mod,
                    void* stream) {
    PDK_TRY("mod_gemv", launch_mod_gemv(tsilu, wmod, bmod, mod, (int)B, (int)n_mod, S(stream)));
    return 0;
}
int pdk_op_adaln(const float* x, const float* mod, int64_t mod_stride, int64_t mod_off, void* xh, void* xl, int64_t B,
                 int64_t S_pad, int64_t c, float eps, void* stream) {
    PDK_TRY("adaln", launch_adaln(x, mod, (int)mod_stride, (int)mod_off, H(xh), H(xl), (int)B, (int)S_pad, (int)c, eps, S(stream)));
    return 0;
}
int pdk_op_precond_adaln(const float* x_hat, const float* coef, const float* a, const float* wx, const float* bx, float* ba,
                         const float* mod, int64_t mod_stride, int64_t mod_off, void* xh, void* xl, int64_t B, int64_t Na,
                         int64_t S_pad, int64_t c_a, float eps, void* stream) {
    PDK_TRY("precond+adaln", launch_precond_adaln(x_hat, coef, 4, a, wx, bx, ba, mod, (int)mod_stride, (int)mod_off, H(xh), H(xl),
                                                  (int)B, (int)Na, (int)S_pad, (int)c_a, eps, S(stream)));
    return 0;
}
int pdk_op_upscale_adaln(float* ba, const float* up, const int32_t* atom2tok, const float* mod, int64_t mod_stride,
                         int64_t mod_off, void* xh, void* xl, int64_t B, int64_t Na, int64_t Sa_pad, int64_t St_pad,
                         int64_t c_a, float eps, void* stream) {
    PDK_TRY("upscale+adaln", launch_upscale_adaln(ba, up, atom2tok, mod, (int)mod_stride, (int)mod_off, H(xh), H(xl), (int)B,
                                                  (int)Na, (int)Sa_pad, (int)St_pad, (int)c_a, eps, S(stream)));
    return 0;
}
int pdk_op_split(const float* x, void* xh, void* xl, int64_t n, void* stream) {
    PDK_TRY("split", launch_split(x, H(xh), H(xl), (size_t)n, S(stream)));
    return 0;
}
static GemmArgs base_args(const void* Ah, const void* Al, int64_t lda, const void* Wh, const void* Wl, int64_t ldw,
                          int64_t M, int64_t N, int64_t K) {
    GemmArgs g{};
    g.Ah = H(Ah); g.Al = H(Al); g.lda = (int)lda;
    g.Wh = H(Wh); g.Wl = H(Wl); g.ldw = (int)ldw;
    g.M = (int)M; g.N = (int)N; g.K = (int)K;
    return g;
}
int pdk_op_gemm_store(const void* Ah, const void* Al, int64_t lda, const void* Wh, const void* Wl, int64_t ldw,
                      int64_t M, int64_t N, int64_t K, const float* bias, int act_silu, float* out, int64_t ldo,
                      void* stream) {
    GemmArgs g = base_args(Ah, Al, lda, Wh, Wl, ldw, M, N, K);
    g.bias = bias; g.act_silu = act_silu; g.out = out; g.ldo = (int)ldo;
    PDK_TRY("gemm_store", launch_gemm(EPI_STORE, g, S(stream)));
    return 0;
}
int pdk_op_gemm_gate_resid(const void* Ah, const void* Al, int64_t lda, const void* Wh, const void* Wl, int64_t ldw,
                           int64_t M, int64_t N, int64_t K, const float* bias, const float* gate, int64_t gate_stride,
                           int64_t rows_per_sample, float* x, int64_t ldx, void* stream) {
    GemmArgs g = base_args(Ah, Al, lda, Wh, Wl, ldw, M, N, K);
    g.bias = bias; g.gate = gate; g.gate_stride = (int)gate_stride; g.rows_per_sample = (int)rows_per_sample;
    g.out = x; g.ldo = (int)ldx;
    PDK_TRY("gemm_gate_resid", launch_gemm(EPI_GATE_RESID, g, S(stream)));
    return 0;
}
int pdk_op_gemm_swiglu(const void* Ah, const void* Al, int64_t lda, const void* Wh, const void* Wl, int64_t ldw,
                       int64_t M, int64_t N, int64_t K, void* ph, void* pl, int64_t ldp, void* stream) {
    GemmArgs g = base_args(Ah, Al, lda, Wh, Wl, ldw, M, N, K);
    g.ph = H(ph); g.pl = H(pl); g.ldp = (int)ldp;
    PDK_TRY("gemm_swiglu", launch_gemm(EPI_SWIGLU, g, S(stream)));
    return 0;
}
int pdk_op_transition_fused(float* x, const float* mod, int64_t mod_stride, int64_t mod_off, const void* w13h,
                            const void* w13l, const void* w2h, const void* w2l, int64_t M, int64_t hidden,
                            int64_t rows_per_sample, float eps, void* stream) {
    if (!x || !mod || !w13h || !w13l || !w2h || !w2l) return fail_msg("pdk_op_transition_fused", "null argument");
    TransitionArgs ta{};
    ta.x = x; ta.mod = mod; ta.mod_stride = (int)mod_stride; ta.mod_off = (int)mod_off;
    ta.w13h = H(w13h); ta.w13l = H(w13l); ta.w2h = H(w2h); ta.w2l = H(w2l);
    ta.M = (int)M; ta.hidden = (int)hidden; ta.rows_per_sample = (int)rows_per_sample; ta.eps = eps;
    PDK_TRY("transition_fused", launch_transition_fused(ta, S(stream)));
    return 0;
}
int pdk_op_gemm_qkv(const void* Ah, const void* Al, int64_t lda, const void* Wh, const void* Wl, int64_t ldw, int64_t M,
                    int64_t c, const float* norm_q, const float* norm_k, float rms_eps, float q_scale,
                    int64_t rows_per_sample, void* q, void* k, void* v, void* stream) {
    GemmArgs g = base_args(Ah, Al, lda, Wh, Wl, ldw, M, 3 * c, c);
    g.norm_q = norm_q; g.norm_k = norm_k; g.c = (int)c; g.rms_eps = rms_eps; g.q_scale = q_scale;
    g.rows_per_sample = (int)rows_per_sample;
    g.q = H(q); g.k = H(k); g.v = H(v);
    PDK_TRY("gemm_qkv", launch_gemm(EPI_QKV, g, S(stream)));
    return 0;
}
int pdk_op_attention(const void* q, const void* k, const void* v, const float* bias, void* oh, void* ol, int64_t B,
                     int64_t Hh, int64_t S_pad, void* stream) {
    AttnArgs a{H(q), H(k), H(v), bias, H(oh), H(ol), (int)B, (int)Hh, (int)S_pad, (int)(Hh * kHeadDim), 0, nullptr};
    PDK_TRY("attention", launch_attention(a, S(stream)));
    return 0;
}
int pdk_op_precond(const float* x_hat, const float* coef, const float* a, const float* wx, const float* bx, float* ba,
                   int64_t B, int64_t Na, int64_t S_pad, int64_t c_a, void* stream) {
    PDK_TRY("precond", launch_precond(x_hat, coef, 4, a, wx, bx, ba, (int)B, (int)Na, (int)S_pad, (int)c_a, S(stream)));
    return 0;
}
int pdk_op_segment_mean(const float* h, const int32_t* tok_start, const float* s, float* bs, int64_t B, int64_t Nt,
                        int64_t Sa_pad, int64_t St_pad, int64_t c_s, void* stream) {
    PDK_TRY("segment_mean", launch_segment_mean(h, tok_start, s, bs, (int)B, (int)Nt, (int)Sa_pad, (int)St_pad, (int)c_s, S(stream)));
    return 0;
}
int pdk_op_gather_add(float* ba, const float* up, const int32_t* atom2tok, int64_t B, int64_t Na, int64_t Sa_pad,
                      int64_t St_pad, int64_t c_a, void* stream) {
    PDK_TRY("gather_add", launch_gather_add(ba, up, atom2tok, (int)B, (int)Na, (int)Sa_pad, (int)St_pad, (int)c_a, S(stream)));
    return 0;
}
int pdk_op_denoise_out(const float* ba, const float* x_hat, const float* coef, const float* ln_w, const float* ln_b,
                       const float* wr, float* x_den, int64_t B, int64_t Na, int64_t S_pad, int64_t c_a, float eps,
                       void* stream) {
    PDK_TRY("denoise_out", launch_denoise_out(ba, x_hat, coef, 4, ln_w, ln_b, wr, x_den, (int)B, (int)Na, (int)S_pad, (int)c_a, eps, nullptr, S(stream)));
    return 0;
}

}  // extern "C"

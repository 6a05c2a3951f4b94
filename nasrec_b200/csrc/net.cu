// Native step executor for the weight-sharing supernet (host code; launches the library's own kernels).
//
// The Python engine (nasrec_b200/engine.py + supernet/*.py) issues ~70 operator calls per training step
// and is host-bound at B=512 (3 ms of interpreter time against 2.2 ms of device work).  This file is the
// same launch sequence -- same operators, same order, same arguments, hence bit-identical results --
// driven from C++: the model is described once (parameter table + per-block node tables), a step takes
// the sampled choice as a flat int array, activations and gradients live in caller-provided arenas
// (bump allocation, no allocator calls), and the optimizer runs on the list of parameters that
// received a gradient.  Reference semantics implemented: SuperNet.forward (supernet.py:513-603),
// SuperNetBlock.forward (:1067-1162), the modules of supernet/modules.py in weight-sharing mode, and the
// step body of train_utils.py:262-286.  Fixed (standalone) models stay on the Python engine / CUDA graphs.
#include <algorithm>
#include <deque>
#include <functional>
#include <map>
#include <new>
#include <stdexcept>
#include <vector>
#include <cstdlib>
#include "common.cuh"

namespace {

constexpr int E = NASREC_EMB_DIM;
constexpr int GROUPS = 8;            // DS_INTERACT_NUM_SPLITS (supernet.py:49)
constexpr float LN_EPS = 1e-5f;
const int BIG_REDUCE_ROWS = getenv("NASREC_BIG_REDUCE_ROWS") ? atoi(getenv("NASREC_BIG_REDUCE_ROWS")) : 2048;   // above this many ids per table the multi-CTA sorted-row reduction takes over (measured crossover, tools/sort_reduce_prof.py: 121 vs 134 us at 2048, 235 vs 164 us at 4096; up to it one CTA per table, radix sort from 1024)

struct OutOfArena : std::runtime_error { OutOfArena() : std::runtime_error("arena") {} };
struct CallFailed : std::runtime_error { int rc; explicit CallFailed(int r) : std::runtime_error("call"), rc(r) {} };
long long g_launches = 0;            // kernels launched through this executor (nominal per entry point, as nasrec_b200/_lib.py)
inline void ck(int rc, int kernels = 1) {
    if (rc) throw CallFailed(rc);
    g_launches += kernels;
}

struct Arena {
    char* base = nullptr;
    size_t cap = 0, off = 0, high = 0;
    float* alloc(int64_t nfloats) {
        const size_t bytes = ((size_t)(nfloats > 0 ? nfloats : 1) * 4 + 255) & ~(size_t)255;
        if (!base || off + bytes > cap) {
            if (off + bytes > high) high = off + bytes;     // tells the caller how much this step needs at least
            throw OutOfArena();
        }
        float* p = (float*)(base + off);
        off += bytes;
        if (off > high) high = off;
        return p;
    }
    void* alloc_bytes(size_t bytes) { return alloc((int64_t)((bytes + 3) / 4)); }
};

struct Var { float* t = nullptr; float* g = nullptr; bool req = false; int64_t n = 0; };
struct Par { float* w; float* st; float* g; int64_t n; int rows, cols; bool req; float* hi = nullptr; float* lo = nullptr; int64_t ldp = 0; int first = 0; };
struct Seg { Var* v; int64_t off; int64_t ld; int64_t width; int64_t w_off; };
using Segs = std::vector<Seg>;

enum NodeType { N_FC = 0, N_DP = 1, N_SUM = 2, N_GATE = 3, N_EFC = 4, N_TRANS = 5, N_ZERO2 = 6, N_ZERO3 = 7 };

struct NodeDesc { int type; int p[24]; };
struct BlockDesc {
    int num_nodes;
    NodeDesc nodes[8];
    int merger[4];      // project_emb_dim: W, b, ln_g, ln_b
    int fm[4];          // deep_fm: W, b, ln_g, ln_b
    int maxd, maxs, dp_P;
};

struct Net {
    int num_blocks = 0, nd = 0, F = 0, final_w = -1, final_b = -1;
    std::vector<Par> par;
    std::vector<BlockDesc> blocks;
    std::vector<int> emb_par;                 // parameter index of each embedding table
    const float* const* d_tables = nullptr;   // device arrays owned by the caller
    const int64_t* d_rows = nullptr;
    float* const* d_table_ptrs_rw = nullptr;
    float* const* d_state_ptrs = nullptr;
    int* d_err = nullptr;
    Arena act, pg;
    size_t pg_dirty = 0;                      // bytes of the parameter-gradient bucket to clear before the next step
    // per-step state
    std::deque<Var> vars;
    std::vector<std::function<void()>> tape;
    bool tape_on = false;
    cudaStream_t st = nullptr;
    std::vector<int> touched;
    std::vector<int> ref_order;               // parameters in the order the forward pass first referenced them
    std::vector<int> ref_stamp;               // (the order the Python engine reduces the gradient norm in)
    int step_id = 0;
    const int64_t* cat_x = nullptr;
    float* emb_gout = nullptr;
    int B = 0;
    // sparse reduction of the step
    int64_t* uniq = nullptr; int* nuniq = nullptr; float* row_grad = nullptr; float* sumsq = nullptr; int sB = 0;
    bool have_sparse = false;
    bool overlap = false;
    bool late_join = false;    // see nasrec_net_set_overlap
    cudaEvent_t join_ev = nullptr;
    bool join_pending = false; // side-stream work of the last backward not yet joined (done by nasrec_net_apply / _grad_bucket)
    bool defer_wgrad = true;   // dense weight gradients queue up during backward and run as one batched launch (nasrec_wgrad_flush)
    // scratch of nasrec_net_sparse_reduce / nasrec_net_apply, reserved by forward_backward so that those two calls cannot
    // overflow an arena after the step's gradients exist (growing the arenas then would leave them pointing at freed memory)
    int reserve_rows = 0;                     // rows of the (all-gathered) batch the sparse reduction will see
    int res_rows = 0;
    int64_t* res_uniq = nullptr; int* res_nuniq = nullptr; float* res_row_grad = nullptr; float* res_sumsq = nullptr; int* res_scratch = nullptr;
    float* res_partial = nullptr; int64_t res_partial_n = 0;
    void* res_big_ws = nullptr; int64_t res_big_bytes = 0;     // workspace of the multi-CTA reduction (large batches)
    bool step_valid = false;                  // false after nasrec_net_set_arenas: gradients of an earlier step are gone
    // data-parallel overlap: called during backward whenever a block's parameter gradients are final, with the
    // byte range of the gradient bucket that was sealed (the caller all-reduces it while backward continues)
    void (*seal_cb)(int64_t offset_bytes, int64_t nbytes) = nullptr;
    size_t seal_min_bytes = 12u << 20;
    std::vector<size_t> block_mark;           // tape length at the start of each block's forward

    Var* var(int64_t n, bool alloc = true) {
        vars.emplace_back();
        Var* v = &vars.back();
        v->n = n;
        if (alloc) v->t = act.alloc(n);
        return v;
    }
    float* grad_of(Var* v) {            // first toucher overwrites (engine.Var.grad_buf)
        if (!v->g) v->g = act.alloc(v->n);
        return v->g;
    }
    float* pgrad(int pi) {
        Par& p = par[pi];
        if (!p.g) {
            p.g = pg.alloc(p.n);        // bucket is cleared at step start, so partial-support writes see zeros
            touched.push_back(pi);
        }
        return p.g;
    }
    void ref(int pi) {
        if (pi < 0) return;
        if (ref_stamp.size() != par.size()) ref_stamp.assign(par.size(), -1);
        if (ref_stamp[pi] != step_id) { ref_stamp[pi] = step_id; ref_order.push_back(pi); }
    }
    void zero(float* p, int64_t n) { cudaMemsetAsync(p, 0, (size_t)n * 4, st); }
    // events for forking a block's sparse node onto the library's side stream during forward
    cudaEvent_t ev[32];
    int ev_made = 0, ev_next = 0;
    cudaEvent_t event() {
        if (ev_made < 32) {
            cudaEventCreateWithFlags(&ev[ev_made], cudaEventDisableTiming);
            return ev[ev_made++];
        }
        ev_next = (ev_next + 1) % 32;
        return ev[ev_next];
    }
    void record(std::function<void()> fn) { if (tape_on) tape.push_back(std::move(fn)); }
};

inline bool preq(Net& n, int pi) { return pi >= 0 && n.par[pi].req; }
// announce the weight's pre-split planes to the GEMM entry points called next (no planes: clears the hint)
inline void hint_planes(const Par& W) {
    if (W.hi) nasrec_set_weight_planes(W.w, W.hi, W.lo, W.ldp, W.rows, W.cols, W.first);
    else nasrec_set_weight_planes(nullptr, nullptr, nullptr, 0, 0, 0, 0);
}
inline float* pw(Net& n, int pi) { return pi >= 0 ? n.par[pi].w : nullptr; }

inline bool any_req(const Segs& s) { for (auto& x : s) if (x.v->req) return true; return false; }

void pack(const Segs& s, nasrec_seg_t* out, bool grad = false) {
    for (size_t i = 0; i < s.size(); ++i) {
        const float* base = grad ? s[i].v->g : s[i].v->t;
        out[i] = nasrec_seg_t{base + s[i].off, s[i].ld, s[i].width, s[i].w_off};
    }
}

// engine._unique_woff_groups
std::vector<Segs> unique_woff_groups(const Segs& list) {
    std::vector<Segs> groups;
    for (auto& s : list) {
        if (s.width == 0) continue;
        bool placed = false;
        for (auto& g : groups) {
            bool clash = false;
            for (auto& o : g) if (o.w_off == s.w_off) { clash = true; break; }
            if (!clash) { g.push_back(s); placed = true; break; }
        }
        if (!placed) groups.push_back(Segs{s});
    }
    return groups;
}

bool distinct_woffs(const Segs& list) {
    for (size_t i = 0; i < list.size(); ++i) {
        if (list[i].width == 0) continue;
        for (size_t j = i + 1; j < list.size(); ++j)
            if (list[j].width > 0 && list[j].w_off == list[i].w_off) return false;
    }
    return true;
}

// engine._grad_targets: false when two segments share a target (grouped general path instead)
bool grad_targets(Net& n, const Segs& list, nasrec_seg_t* dsegs, int* flags) {
    for (size_t i = 0; i < list.size(); ++i) {
        if (!(list[i].v->req && list[i].width > 0)) continue;
        for (size_t j = i + 1; j < list.size(); ++j)
            if (list[j].v->req && list[j].width > 0 && list[j].v == list[i].v && list[j].off == list[i].off) return false;
    }
    std::vector<Var*> fresh;
    for (size_t i = 0; i < list.size(); ++i) {
        const Seg& s = list[i];
        if (s.v->req && s.width > 0) {
            if (!s.v->g) { s.v->g = n.act.alloc(s.v->n); fresh.push_back(s.v); }
            bool is_fresh = false;
            for (Var* f : fresh) if (f == s.v) is_fresh = true;
            flags[i] = is_fresh ? 0 : 1;
            dsegs[i] = nasrec_seg_t{s.v->g + s.off, s.ld, s.width, s.w_off};
        } else {
            flags[i] = 0;
            dsegs[i] = nasrec_seg_t{nullptr, s.ld, s.width, s.w_off};
        }
    }
    return true;
}

// engine._dgrad_groups
std::vector<std::pair<Segs, int>> dgrad_groups(Net& n, const Segs& list) {
    Segs fresh;
    std::vector<Segs> acc;
    std::vector<Var*> seen_fresh;
    for (auto& s : list) {
        if (!s.v->req || s.width == 0) continue;
        bool in_seen = false;
        for (Var* f : seen_fresh) if (f == s.v) in_seen = true;
        bool key_in_fresh = false;
        for (auto& o : fresh) if (o.v == s.v && o.off == s.off) key_in_fresh = true;
        if (!s.v->g || (in_seen && !key_in_fresh)) {
            if (!s.v->g) { s.v->g = n.act.alloc(s.v->n); seen_fresh.push_back(s.v); }
            fresh.push_back(s);
        } else {
            bool placed = false;
            for (auto& g : acc) {
                bool clash = false;
                for (auto& o : g) if (o.v == s.v && o.off == s.off) { clash = true; break; }
                if (!clash) { g.push_back(s); placed = true; break; }
            }
            if (!placed) acc.push_back(Segs{s});
        }
    }
    std::vector<std::pair<Segs, int>> out;
    if (!fresh.empty()) out.emplace_back(fresh, 0);
    for (auto& g : acc) out.emplace_back(g, 1);
    return out;
}

void accumulate_into(Net& n, Var* v, float* d) {       // engine._accumulate_into
    if (!v->g) { v->g = d; return; }
    ck(nasrec_act_fwd(d, v->n, 1, (int)v->n, 0, v->g, v->n, 1, n.st));
}

// ------------------------------------------------------------------------------------------- linear (+LN/act)
struct LinArgs {
    int W = -1, b = -1, lng = -1, lnb = -1;
    bool relu = false;
    int d_out = 0, n_off = 0, n_full = -1;
    Var* out = nullptr;
    int64_t out_off = 0, ldy = -1;
    int accumulate = 0;
    bool ln_first = false;       // FactorizationMachine3D looks its LayerNorm up before its weight
};

Var* linear_ln(Net& n, const Segs& segs, int M, LinArgs a) {
    if (a.ln_first) { n.ref(a.lng); n.ref(a.lnb); }
    n.ref(a.W); n.ref(a.b); n.ref(a.lng); n.ref(a.lnb);
    Par& W = n.par[a.W];
    const int64_t ldw = W.cols;
    const int n_full = a.n_full < 0 ? W.rows - a.n_off : a.n_full;
    const bool has_ln = a.lng >= 0;
    const int N = has_ln ? n_full : a.d_out;
    float* z = n.act.alloc((int64_t)M * N);
    const int ns = (int)segs.size();
    nasrec_seg_t sp[NASREC_MAX_SEGS];
    if (ns < 1 || ns > NASREC_MAX_SEGS) throw CallFailed(NASREC_EINVAL);
    pack(segs, sp);
    Var* out = a.out;
    int64_t ldy = a.ldy;
    if (!out) { out = n.var((int64_t)M * a.d_out); ldy = a.d_out; }
    float *mean = nullptr, *rstd = nullptr;
    if (has_ln) { mean = n.act.alloc(M); rstd = n.act.alloc(M); }
    hint_planes(W);
    ck(nasrec_linear_ln_fwd(sp, ns, W.w, ldw, a.n_off, N, pw(n, a.b), pw(n, a.lng), pw(n, a.lnb), LN_EPS, a.relu, a.d_out, z,
                            out->t + a.out_off, ldy, mean, rstd, a.accumulate, M, n.st), 2);
    const bool req = any_req(segs) || W.req || preq(n, a.b) || preq(n, a.lng) || preq(n, a.lnb);
    out->req = out->req || req;
    if (!(n.tape_on && req)) return out;
    Net* np = &n;
    n.record([np, segs, M, a, out, ldy, z, mean, rstd, N, ldw]() {
        Net& n = *np;
        if (!out->g) return;
        Par& W = n.par[a.W];
        const int ns = (int)segs.size();
        float* dz = n.act.alloc((int64_t)M * N);
        const bool has_ln = a.lng >= 0;
        const bool want_ln = has_ln && (preq(n, a.lng) || preq(n, a.lnb));
        float* gw = W.req ? n.pgrad(a.W) : nullptr;
        float* gb = preq(n, a.b) ? n.pgrad(a.b) : nullptr;
        nasrec_seg_t sp[NASREC_MAX_SEGS], dsp[NASREC_MAX_SEGS];
        int flags[NASREC_MAX_SEGS];
        pack(segs, sp);
        hint_planes(W);
        if (distinct_woffs(segs) && grad_targets(n, segs, dsp, flags)) {
            ck(nasrec_linear_ln_bwd(out->g + a.out_off, ldy, a.d_out, z, M, N, pw(n, a.lng), pw(n, a.lnb), mean, rstd, a.relu,
                                    sp, dsp, flags, ns, W.w, ldw, a.n_off, gw, gb, want_ln ? n.pgrad(a.lng) : nullptr,
                                    want_ln ? n.pgrad(a.lnb) : nullptr, dz, n.st), 5);
            return;
        }
        // general path: the same source feeds two segments (Sum / Gating with left == right)
        if (has_ln)
            ck(nasrec_ln_bwd(out->g + a.out_off, ldy, a.d_out, z, N, M, N, pw(n, a.lng), pw(n, a.lnb), mean, rstd, a.relu, dz, N,
                             want_ln ? n.pgrad(a.lng) : nullptr, want_ln ? n.pgrad(a.lnb) : nullptr, 0, n.st), 2);
        else
            ck(nasrec_act_bwd(out->g + a.out_off, ldy, z, N, M, N, a.relu, dz, N, n.st));
        if (gw) {
            int gi = 0;
            for (auto& grp : unique_woff_groups(segs)) {
                nasrec_seg_t g[NASREC_MAX_SEGS];
                pack(grp, g);
                ck(nasrec_seg_linear_wgrad(dz, N, N, g, (int)grp.size(), gw, ldw, a.n_off, M, gi ? 1 : 0, n.st));
                ++gi;
            }
        }
        if (gb) ck(nasrec_colsum(dz, N, M, N, gb + a.n_off, 0, n.st));
        for (auto& ga : dgrad_groups(n, segs)) {
            nasrec_seg_t g[NASREC_MAX_SEGS];
            pack(ga.first, g, true);
            ck(nasrec_seg_linear_dgrad(dz, N, N, W.w, ldw, a.n_off, g, (int)ga.first.size(), M, ga.second, n.st));
        }
    });
    return out;
}

// ------------------------------------------------------------------------------------------- sparse-axis projection
struct SprojArgs {
    int W = -1, b = -1, lng = -1, lnb = -1;
    bool relu = false;
    int p_out = 0;
    Var* out = nullptr;
    int64_t out_off = 0, out_bstride = -1;
    int accumulate = 0;
};

Var* sproj_ln(Net& n, const Segs& segs, int B, SprojArgs a) {
    n.ref(a.W); n.ref(a.b); n.ref(a.lng); n.ref(a.lnb);
    Par& W = n.par[a.W];
    const int64_t ldw = W.cols;
    const int P_full = W.rows;
    const bool has_ln = a.lng >= 0;
    const int P = has_ln ? P_full : a.p_out;
    float* z = n.act.alloc((int64_t)B * P * E);
    const int ns = (int)segs.size();
    if (ns < 1 || ns > NASREC_MAX_SEGS) throw CallFailed(NASREC_EINVAL);
    nasrec_seg_t sp[NASREC_MAX_SEGS];
    pack(segs, sp);
    Var* out = a.out;
    int64_t obs = a.out_bstride;
    if (!out) { out = n.var((int64_t)B * a.p_out * E); obs = (int64_t)a.p_out * E; }
    float *mean = nullptr, *rstd = nullptr;
    if (has_ln) { mean = n.act.alloc((int64_t)B * E); rstd = n.act.alloc((int64_t)B * E); }
    hint_planes(W);
    ck(nasrec_sproj_ln_fwd(sp, ns, W.w, ldw, P, pw(n, a.b), pw(n, a.lng), pw(n, a.lnb), LN_EPS, a.relu, a.p_out, z,
                           out->t + a.out_off, obs, mean, rstd, a.accumulate, B, n.st), 2);
    const bool req = any_req(segs) || W.req || preq(n, a.b) || preq(n, a.lng) || preq(n, a.lnb);
    out->req = out->req || req;
    if (!(n.tape_on && req)) return out;
    Net* np = &n;
    n.record([np, segs, B, a, out, obs, z, mean, rstd, P, ldw]() {
        Net& n = *np;
        if (!out->g) return;
        Par& W = n.par[a.W];
        const int ns = (int)segs.size();
        float* dz = n.act.alloc((int64_t)B * P * E);
        const bool has_ln = a.lng >= 0;
        const bool want_ln = has_ln && (preq(n, a.lng) || preq(n, a.lnb));
        float* gw = W.req ? n.pgrad(a.W) : nullptr;
        float* gb = preq(n, a.b) ? n.pgrad(a.b) : nullptr;
        nasrec_seg_t sp[NASREC_MAX_SEGS], dsp[NASREC_MAX_SEGS];
        int flags[NASREC_MAX_SEGS];
        pack(segs, sp);
        hint_planes(W);
        if (distinct_woffs(segs) && grad_targets(n, segs, dsp, flags)) {
            float* ws = nullptr;
            if (gw) {
                int64_t tw = 0;
                for (auto& s : segs) tw += s.width;
                ws = n.act.alloc(nasrec_sproj_wgrad_ws_floats(P, tw, B));
            }
            ck(nasrec_sproj_ln_bwd(out->g + a.out_off, obs, a.p_out, z, B, P, pw(n, a.lng), pw(n, a.lnb), mean, rstd, a.relu, sp,
                                   dsp, flags, ns, W.w, ldw, gw, gb, want_ln ? n.pgrad(a.lng) : nullptr,
                                   want_ln ? n.pgrad(a.lnb) : nullptr, dz, ws, n.st), 6);
            return;
        }
        const int64_t zbs = (int64_t)P * E;
        if (has_ln)
            ck(nasrec_ln3_bwd(out->g + a.out_off, obs, a.p_out, z, zbs, B, P, pw(n, a.lng), pw(n, a.lnb), mean, rstd, a.relu, dz,
                              zbs, want_ln ? n.pgrad(a.lng) : nullptr, want_ln ? n.pgrad(a.lnb) : nullptr, 0, n.st), 2);
        else
            ck(nasrec_act_bwd(out->g + a.out_off, obs, z, zbs, B, P * E, a.relu, dz, zbs, n.st));
        if (gw) {
            int gi = 0;
            for (auto& grp : unique_woff_groups(segs)) {
                int64_t tw = 0;
                for (auto& s : grp) tw += s.width;
                float* ws = n.act.alloc(nasrec_sproj_wgrad_ws_floats(P, tw, B));
                nasrec_seg_t g[NASREC_MAX_SEGS];
                pack(grp, g);
                ck(nasrec_sproj_wgrad(dz, zbs, P, g, (int)grp.size(), gw, ldw, B, gi ? 1 : 0, ws, n.st), 2);
                ++gi;
            }
        }
        if (gb) ck(nasrec_sproj_bias_grad(dz, zbs, P, B, gb, 0, n.st));
        for (auto& ga : dgrad_groups(n, segs)) {
            nasrec_seg_t g[NASREC_MAX_SEGS];
            pack(ga.first, g, true);
            ck(nasrec_sproj_dgrad(dz, zbs, P, W.w, ldw, g, (int)ga.first.size(), B, ga.second, n.st));
        }
    });
    return out;
}

// ------------------------------------------------------------------------------------------- small operators
Var* dot_tril(Net& n, Var* x, Var* y, int B, int P) {
    const int R = (P + 1) * P / 2;
    const int ldr = (R + 3) & ~3;          // 16-byte aligned rows (the consumer GEMM fetches them by TMA)
    Var* out = n.var((int64_t)B * ldr);
    ck(nasrec_dot_tril_fwd(x->t, E, y->t, (int64_t)P * E, P, out->t, ldr, B, n.st));
    out->req = x->req || y->req;
    if (!(n.tape_on && out->req)) return out;
    Net* np = &n;
    n.record([np, x, y, B, P, ldr, out]() {
        Net& n = *np;
        if (!out->g) return;
        float* dx = x->req ? n.act.alloc((int64_t)B * E) : nullptr;
        float* dy = y->req ? n.act.alloc((int64_t)B * P * E) : nullptr;
        ck(nasrec_dot_tril_bwd(out->g, ldr, x->t, E, y->t, (int64_t)P * E, P, dx, E, dy, (int64_t)P * E, B, n.st));
        if (dx) accumulate_into(n, x, dx);
        if (dy) accumulate_into(n, y, dy);
    });
    return out;
}

Var* gate(Net& n, Var* pre, const Segs& right, int M, int K) {
    Var* out = n.var((int64_t)M * K);
    const int ns = (int)right.size();
    nasrec_seg_t sp[NASREC_MAX_SEGS];
    pack(right, sp);
    ck(nasrec_gate_fwd(pre->t, K, sp, ns, out->t, K, M, n.st));
    out->req = pre->req || any_req(right);
    if (!(n.tape_on && out->req)) return out;
    Net* np = &n;
    n.record([np, pre, right, M, K, out]() {
        Net& n = *np;
        if (!out->g) return;
        const int ns = (int)right.size();
        nasrec_seg_t sp[NASREC_MAX_SEGS];
        pack(right, sp);
        float* dpre = n.act.alloc((int64_t)M * K);
        std::vector<int> fresh, acc;
        for (int i = 0; i < ns; ++i) {
            const Seg& s = right[i];
            if (s.v->req && !s.v->g) { s.v->g = n.act.alloc(s.v->n); fresh.push_back(i); }
            else if (s.v->req) acc.push_back(i);
        }
        auto launch = [&](const std::vector<int>& chosen, int accf) {
            nasrec_seg_t dsp[NASREC_MAX_SEGS];
            for (int i = 0; i < ns; ++i) dsp[i] = nasrec_seg_t{nullptr, 0, 0, 0};
            for (int i : chosen) dsp[i] = nasrec_seg_t{right[i].v->g + right[i].off, right[i].ld, right[i].width, right[i].w_off};
            ck(nasrec_gate_bwd(out->g, K, pre->t, K, sp, dsp, ns, dpre, K, M, accf, n.st));
        };
        if (!fresh.empty()) launch(fresh, 0);
        if (!acc.empty()) launch(acc, 1);
        if (fresh.empty() && acc.empty()) launch({}, 0);
        if (pre->req) accumulate_into(n, pre, dpre);
    });
    return out;
}

Var* fm_ix(Net& n, Var* x, int B, int rows, int64_t bstride) {
    Var* out = n.var((int64_t)B * E);
    ck(nasrec_fm_fwd(x->t, bstride, rows, out->t, B, n.st));
    out->req = x->req;
    if (!(n.tape_on && out->req)) return out;
    Net* np = &n;
    n.record([np, x, B, rows, bstride, out]() {
        Net& n = *np;
        if (!out->g) return;
        if (!x->g) {
            x->g = n.act.alloc(x->n);
            if ((int64_t)rows * E < bstride) n.zero(x->g, x->n);
            ck(nasrec_fm_bwd(out->g, x->t, bstride, rows, nullptr, 0, x->g, bstride, B, n.st));
        } else {
            ck(nasrec_fm_bwd(out->g, x->t, bstride, rows, x->g, bstride, x->g, bstride, B, n.st));
        }
    });
    return out;
}

void copy2d(Net& n, Var* src, int64_t src_off, int64_t lds, int M, int N, Var* dst, int64_t dst_off, int64_t ldd, int accumulate) {
    ck(nasrec_act_fwd(src->t + src_off, lds, M, N, 0, dst->t + dst_off, ldd, accumulate, n.st));
    dst->req = dst->req || src->req;
    if (!(n.tape_on && src->req)) return;
    Net* np = &n;
    n.record([np, src, src_off, lds, M, N, dst, dst_off, ldd]() {
        Net& n = *np;
        if (!dst->g) return;
        if (!src->g) {
            const bool full = src_off == 0 && N == lds && src->n == (int64_t)M * N;
            src->g = n.act.alloc(src->n);
            if (!full) n.zero(src->g, src->n);
            ck(nasrec_act_fwd(dst->g + dst_off, ldd, M, N, 0, src->g + src_off, lds, 0, n.st));
        } else {
            ck(nasrec_act_fwd(dst->g + dst_off, ldd, M, N, 0, src->g + src_off, lds, 1, n.st));
        }
    });
}

// attention core; params = 12 parameter indices in the order of include/nasrec_b200.h
Var* attention(Net& n, Var* x, int B, int L, int s_live, const int* params, Var* out, int64_t out_bstride, bool accumulate_out) {
    const float* pp[12];
    bool preq_any = false;
    for (int i = 0; i < 12; ++i) n.ref(params[i]);
    for (int i = 0; i < 12; ++i) { pp[i] = n.par[params[i]].w; preq_any = preq_any || n.par[params[i]].req; }
    if (accumulate_out) {
        float* tmp = n.act.alloc((int64_t)B * s_live * E);
        ck(nasrec_attn_fwd(x->t, (int64_t)s_live * E, L, s_live, pp, tmp, (int64_t)s_live * E, B, n.st));
        ck(nasrec_act_fwd(tmp, (int64_t)s_live * E, B, s_live * E, 0, out->t, out_bstride, 1, n.st));
    } else {
        ck(nasrec_attn_fwd(x->t, (int64_t)s_live * E, L, s_live, pp, out->t, out_bstride, B, n.st));
    }
    const bool req = x->req || preq_any;
    out->req = out->req || req;
    if (!(n.tape_on && req)) return out;
    Net* np = &n;
    std::vector<int> pidx(params, params + 12);
    n.record([np, x, B, L, s_live, pidx, out, out_bstride, preq_any]() {
        Net& n = *np;
        if (!out->g) return;
        const float* pp[12];
        for (int i = 0; i < 12; ++i) pp[i] = n.par[pidx[i]].w;
        float* dx = x->req ? n.act.alloc((int64_t)B * s_live * E) : nullptr;
        float* dpar = nullptr;
        if (preq_any) {
            dpar = n.pg.alloc(NASREC_ATTN_PARAMS);       // bucket is pre-cleared; kernel overwrites (accumulate 0)
            int64_t o = 0;
            for (int i = 0; i < 12; ++i) {
                Par& p = n.par[pidx[i]];
                if (p.req) { p.g = dpar + o; n.touched.push_back(pidx[i]); }
                o += p.n;
            }
        }
        float* ws = n.act.alloc(nasrec_attn_bwd_ws_floats(B));
        ck(nasrec_attn_bwd(out->g, out_bstride, x->t, (int64_t)s_live * E, L, s_live, pp, dx, (int64_t)s_live * E, dpar, 0, ws, B,
                           n.st), 2);
        if (dx) accumulate_into(n, x, dx);
    });
    return out;
}

// ------------------------------------------------------------------------------------------- choice decoding
struct BlockChoice {
    std::vector<int> dense, sparse, left, right, active;
    int d = 0, s = 0, dsi = 0, dfm = 0;
};

// flat layout per block: [n_dense, idx x8, n_sparse, idx x8, n_left, idx x8, n_right, idx x8, n_active, idx x8, d, s, dsi, dfm]
constexpr int CHOICE_STRIDE = 5 * 9 + 4;

void sorted_unique(std::vector<int>& v) {
    std::sort(v.begin(), v.end());
    v.erase(std::unique(v.begin(), v.end()), v.end());
}

BlockChoice decode(const int* c) {
    BlockChoice bc;
    std::vector<int>* lists[5] = {&bc.dense, &bc.sparse, &bc.left, &bc.right, &bc.active};
    for (int l = 0; l < 5; ++l) {
        const int cnt = c[l * 9];
        if (cnt < 0 || cnt > 8) throw CallFailed(NASREC_EINVAL);
        for (int i = 0; i < cnt; ++i) lists[l]->push_back(c[l * 9 + 1 + i]);
    }
    for (int l = 0; l < 4; ++l) sorted_unique(*lists[l]);      // supernet._segments: sorted(set(...))
    bc.d = c[45]; bc.s = c[46]; bc.dsi = c[47]; bc.dfm = c[48];
    return bc;
}

struct DSrc { Var* v; int w; int ld; };
struct SSrc { Var* v; int s, g; };

bool is_dense_binary(int t) { return t == N_SUM || t == N_GATE; }

}  // namespace

// ------------------------------------------------------------------------------------------- the network
namespace {

void run_block(Net& n, int bi, const BlockChoice& ch, const std::vector<DSrc>& dsrc, const std::vector<SSrc>& ssrc,
               const std::vector<bool>& have, int B, DSrc& d_out, SSrc& s_out) {
    const BlockDesc& bd = n.blocks[bi];
    const int maxd = bd.maxd, maxs = bd.maxs;
    const int n_src = (int)dsrc.size();
    const int d = ch.d, s = ch.s;
    if (ch.dsi != 0 && ch.dsi != 1) throw CallFailed(NASREC_EINVAL);
    if (d > maxd || s > maxs || d <= 0 || s <= 0) throw CallFailed(NASREC_EINVAL);
    const int g = ch.dsi == 1 ? GROUPS : 0;
    bool need_dense = false, need_sparse = false, need_lr = false;
    for (int a : ch.active) {
        if (a < 0 || a >= bd.num_nodes) throw CallFailed(NASREC_EINVAL);
        const int t = bd.nodes[a].type;
        if (t == N_FC || t == N_DP) need_dense = true;
        if (t == N_EFC || t == N_TRANS || t == N_DP) need_sparse = true;
        if (is_dense_binary(t)) need_lr = true;
    }
    auto dense_segs = [&](const std::vector<int>& idx, bool want, int& total) {
        Segs segs;
        total = 0;
        if (!want) return segs;
        for (int j : idx) {
            if (j < 0 || j >= n_src || !have[j]) throw CallFailed(NASREC_EINVAL);
            const int64_t w_off = j == 0 ? 0 : n.nd + (int64_t)maxd * (j - 1);
            segs.push_back(Seg{dsrc[j].v, 0, dsrc[j].ld, dsrc[j].w, w_off});
        }
        total = n_src > 1 ? n.nd + maxd * (n_src - 1) : n.nd;
        return segs;
    };
    int Kd = 0, Ks = 0, Kl = 0, Kr = 0;
    Segs dsegs = dense_segs(ch.dense, need_dense, Kd);
    Segs lsegs = dense_segs(ch.left, need_lr, Kl);
    Segs rsegs = dense_segs(ch.right, need_lr, Kr);
    Segs ssegs;
    if (need_sparse) {
        for (int j : ch.sparse) {
            if (j < 0 || j >= n_src || !have[j]) throw CallFailed(NASREC_EINVAL);
            const SSrc& src = ssrc[j];
            const int64_t bs = (int64_t)(src.s + src.g) * E;
            const int64_t base = j == 0 ? 0 : n.F + (int64_t)(maxs + GROUPS) * (j - 1);
            ssegs.push_back(Seg{src.v, 0, bs, src.s, base});
            if (src.g) ssegs.push_back(Seg{src.v, (int64_t)src.s * E, bs, src.g, base + maxs});
        }
        Ks = n_src > 1 ? n.F + (maxs + GROUPS) * (n_src - 1) : n.F;
    }
    (void)Kd; (void)Ks; (void)Kl; (void)Kr;

    Var* dense_out = n.var((int64_t)B * d);
    const int rows = s + g;
    Var* sparse_out = n.var((int64_t)B * rows * E);
    int nd_w = 0, ns_w = 0;
    // The dense node and the sparse node of a block are independent until the merger: with a side stream attached
    // the sparse node's forward kernels (sparse-axis GEMM, LayerNorm, attention) run there, concurrently with the
    // dense node's, and are joined before the merger / FM read sparse_out.
    cudaStream_t main_st = n.st;
    cudaStream_t side = n.overlap ? nasrec_internal_side_stream() : nullptr;
    if (side == main_st) side = nullptr;
    bool forked = false;
    auto on_side = [&]() {
        if (!side) return;
        if (!forked) {
            cudaEvent_t e = n.event();
            cudaEventRecord(e, main_st);
            cudaStreamWaitEvent(side, e, 0);
            forked = true;
        }
        n.st = side;
    };
    for (int ai : ch.active) {
        const NodeDesc& nd = bd.nodes[ai];
        const int* p = nd.p;
        n.st = main_st;
        if (nd.type == N_EFC || nd.type == N_TRANS) on_side();
        switch (nd.type) {
        case N_FC: {
            LinArgs a; a.W = p[0]; a.b = p[1]; a.lng = p[2]; a.lnb = p[3]; a.relu = true; a.d_out = d;
            a.out = dense_out; a.ldy = d; a.accumulate = nd_w > 0;
            linear_ln(n, dsegs, B, a);
            ++nd_w;
        } break;
        case N_DP: {
            const int P = bd.dp_P;
            LinArgs a; a.W = p[0]; a.b = p[1]; a.lng = p[2]; a.lnb = p[3]; a.d_out = E;
            Var* x = linear_ln(n, dsegs, B, a);
            SprojArgs sa; sa.W = p[4]; sa.b = p[5]; sa.lng = p[6]; sa.lnb = p[7]; sa.p_out = P;
            Var* y = sproj_ln(n, ssegs, B, sa);
            Var* R = dot_tril(n, x, y, B, P);
            const int nR = (P + 1) * P / 2;
            LinArgs o; o.W = p[8]; o.b = p[9]; o.lng = p[10]; o.lnb = p[11]; o.d_out = d;
            o.out = dense_out; o.ldy = d; o.accumulate = nd_w > 0;
            linear_ln(n, Segs{Seg{R, 0, (nR + 3) & ~3, nR, 0}}, B, o);
            ++nd_w;
        } break;
        case N_SUM: {
            Segs both = lsegs;
            both.insert(both.end(), rsegs.begin(), rsegs.end());
            LinArgs a; a.W = p[0]; a.b = p[1]; a.lng = p[2]; a.lnb = p[3]; a.d_out = d;
            a.out = dense_out; a.ldy = d; a.accumulate = nd_w > 0;
            linear_ln(n, both, B, a);
            ++nd_w;
        } break;
        case N_GATE: {
            Segs live;
            int Kg = 0;
            for (auto& sg : rsegs) if (sg.width > 0) { live.push_back(sg); Kg += (int)sg.width; }
            Var* pre = n.var((int64_t)B * Kg);
            int ko = 0;
            for (auto& sg : live) {
                LinArgs a; a.W = p[0]; a.b = p[1]; a.d_out = (int)sg.width; a.n_off = (int)sg.w_off; a.n_full = (int)sg.width;
                a.out = pre; a.out_off = ko; a.ldy = Kg;
                linear_ln(n, lsegs, B, a);
                ko += (int)sg.width;
            }
            Var* gated = gate(n, pre, live, B, Kg);
            Segs gsegs;
            ko = 0;
            for (auto& sg : live) { gsegs.push_back(Seg{gated, ko, Kg, sg.width, sg.w_off}); ko += (int)sg.width; }
            LinArgs o; o.W = p[2]; o.b = p[3]; o.lng = p[4]; o.lnb = p[5]; o.d_out = d;
            o.out = dense_out; o.ldy = d; o.accumulate = nd_w > 0;
            linear_ln(n, gsegs, B, o);
            ++nd_w;
        } break;
        case N_EFC: {
            SprojArgs a; a.W = p[0]; a.b = p[1]; a.lng = p[2]; a.lnb = p[3]; a.relu = true; a.p_out = s;
            a.out = sparse_out; a.out_bstride = (int64_t)rows * E; a.accumulate = ns_w > 0;
            sproj_ln(n, ssegs, B, a);
            ++ns_w;
        } break;
        case N_TRANS: {
            SprojArgs a; a.W = p[0]; a.b = p[1]; a.lng = p[2]; a.lnb = p[3]; a.p_out = s;
            Var* xa = sproj_ln(n, ssegs, B, a);
            attention(n, xa, B, maxs, s, p + 4, sparse_out, (int64_t)rows * E, ns_w > 0);
            ++ns_w;
        } break;
        case N_ZERO2: case N_ZERO3: break;
        default: throw CallFailed(NASREC_EINVAL);
        }
    }
    n.st = main_st;
    if (forked) {
        cudaEvent_t e = n.event();
        cudaEventRecord(e, side);
        cudaStreamWaitEvent(main_st, e, 0);
    }
    if (nd_w == 0) n.zero(dense_out->t, dense_out->n);
    if (ns_w == 0) n.zero(sparse_out->t, sparse_out->n);

    // dense -> sparse merger reads the node sum before the FM term is added (supernet.py:1137-1157)
    Var* dense_sum = dense_out;
    const bool project = ch.dsi == 1 && maxd != E * GROUPS;
    if (project) {
        LinArgs a; a.W = bd.merger[0]; a.b = bd.merger[1]; a.lng = bd.merger[2]; a.lnb = bd.merger[3]; a.d_out = GROUPS * E;
        a.out = sparse_out; a.out_off = (int64_t)s * E; a.ldy = (int64_t)rows * E;
        linear_ln(n, Segs{Seg{dense_sum, 0, d, d, 0}}, B, a);
    }
    if (ch.dfm == 1) {
        if (project && n.tape_on) {
            dense_out = n.var((int64_t)B * d);
            copy2d(n, dense_sum, 0, d, B, d, dense_out, 0, d, 0);
        }
        Var* ix = fm_ix(n, sparse_out, B, s, (int64_t)rows * E);      // FM sees the node rows only, not the merger rows
        if (bd.fm[0] < 0) throw CallFailed(NASREC_EINVAL);     // max_dims == 16 corner stays on the Python engine
        LinArgs a; a.W = bd.fm[0]; a.b = bd.fm[1]; a.lng = bd.fm[2]; a.lnb = bd.fm[3]; a.d_out = d;
        a.out = dense_out; a.ldy = d; a.accumulate = 1; a.ln_first = true;
        linear_ln(n, Segs{Seg{ix, 0, E, E, 0}}, B, a);
    }
    if (ch.dsi == 1 && !project) copy2d(n, dense_out, 0, d, B, d, sparse_out, (int64_t)s * E, (int64_t)rows * E, 0);
    d_out = DSrc{dense_out, d, d};
    s_out = SSrc{sparse_out, s, g};
}

std::vector<bool> liveness(Net& n, const std::vector<BlockChoice>& ch) {
    const int nb = n.num_blocks;
    std::vector<bool> need(nb + 1, false);
    need[nb] = true;
    for (int i = nb - 1; i >= 0; --i) {
        if (!need[i + 1]) continue;
        bool fc_dp = false, bin = false, sp = false;
        for (int a : ch[i].active) {
            const int t = n.blocks[i].nodes[a].type;
            if (t == N_FC || t == N_DP) fc_dp = true;
            if (is_dense_binary(t)) bin = true;
            if (t == N_TRANS || t == N_EFC || t == N_DP) sp = true;
        }
        if (fc_dp) for (int j : ch[i].dense) need[j] = true;
        if (bin) { for (int j : ch[i].left) need[j] = true; for (int j : ch[i].right) need[j] = true; }
        if (sp) for (int j : ch[i].sparse) need[j] = true;
    }
    return need;       // need[j]: source j (0 = stem, i+1 = block i)
}

void reset_step(Net& n, cudaStream_t st) {
    n.vars.clear();
    n.tape.clear();
    n.act.off = 0;
    n.pg.off = 0;
    for (int pi : n.touched) n.par[pi].g = nullptr;
    n.touched.clear();
    n.ref_order.clear();
    n.block_mark.clear();
    ++n.step_id;
    n.have_sparse = false;
    n.emb_gout = nullptr;
    n.st = st;
}

Var* forward(Net& n, const int* choice, const float* int_x, const int64_t* cat_x, const float* emb_rows, int B, bool train) {
    std::vector<BlockChoice> ch;
    for (int i = 0; i < n.num_blocks; ++i) ch.push_back(decode(choice + i * CHOICE_STRIDE));
    for (auto& c : ch) for (int a : c.active) if (a < 0 || a >= 8) throw CallFailed(NASREC_EINVAL);
    n.tape_on = train;
    n.B = B;
    n.cat_x = cat_x;
    // stem
    // the dense input keeps the caller's [B, nd] layout; a row stride that is not a multiple of 4 floats (nd = 13)
    // cannot be described by a tensor map, so it is staged once into a padded copy that TMA can fetch
    const int nd_ld = (n.nd + 3) & ~3;
    Var* x0 = n.var((int64_t)B * nd_ld, nd_ld != n.nd);
    if (nd_ld != n.nd) ck(nasrec_act_fwd(int_x, n.nd, B, n.nd, 0, x0->t, nd_ld, 0, n.st));
    else x0->t = const_cast<float*>(int_x);
    bool emb_req = false;
    for (int pi : n.emb_par) emb_req = emb_req || n.par[pi].req;
    Var* sp0 = n.var((int64_t)B * n.F * E, emb_rows == nullptr);
    if (emb_rows) sp0->t = const_cast<float*>(emb_rows);       // frozen tables: rows gathered once per batch by the caller
    else ck(nasrec_emb_gather_fwd(n.d_tables, n.d_rows, cat_x, sp0->t, B, n.F, n.d_err, n.st));
    sp0->req = train && emb_req;
    Var* stem = sp0;
    if (sp0->req) {
        Net* np = &n;
        n.record([np, stem]() { np->emb_gout = stem->g; });
    }
    std::vector<DSrc> dsrc{DSrc{x0, n.nd, nd_ld}};
    std::vector<SSrc> ssrc{SSrc{sp0, n.F, 0}};
    std::vector<bool> have{true};
    const std::vector<bool> need = liveness(n, ch);
    for (int i = 0; i < n.num_blocks; ++i) {
        if (!need[i + 1]) {
            dsrc.push_back(DSrc{nullptr, 0, 0});
            ssrc.push_back(SSrc{nullptr, 0, 0});
            have.push_back(false);
            continue;
        }
        DSrc d;
        SSrc s;
        n.block_mark.push_back(n.tape.size());
        run_block(n, i, ch[i], dsrc, ssrc, have, B, d, s);
        dsrc.push_back(d);
        ssrc.push_back(s);
        have.push_back(true);
    }
    const DSrc& dl = dsrc.back();
    const SSrc& sl = ssrc.back();
    const BlockDesc& last = n.blocks[n.num_blocks - 1];
    const int64_t bs = (int64_t)(sl.s + sl.g) * E;
    Segs segs{Seg{dl.v, 0, dl.ld, dl.w, 0}, Seg{sl.v, 0, bs, (int64_t)sl.s * E, last.maxd}};
    if (sl.g) segs.push_back(Seg{sl.v, (int64_t)sl.s * E, bs, (int64_t)sl.g * E, last.maxd + (int64_t)last.maxs * E});
    LinArgs a; a.W = n.final_w; a.b = n.final_b; a.d_out = 1;
    return linear_ln(n, segs, B, a);
}

// ------------------------------------------------------------------------------------------- many candidates, shared work
// One-shot scoring evaluates many subnets against the SAME weights on the SAME batch (searcher_utils.py:57-104,
// eval_subnet_from_supernet.py:182-198).  A block's output is a pure function of its own choice and of the outputs of the
// sources it reads, so candidates that agree on a block and on everything upstream of it share that block: it is computed
// once and its output tensors are handed to every candidate that needs them (the children of one EA generation differ
// from their parent in one field of one block, so on average half of their blocks are shared; every candidate shares the
// stem).  Keys are exact (the full dependency description, no hashing); shared tensors are read-only.
struct MultiStats { int computed = 0, reused = 0; };

std::vector<long long> block_key(Net& n, int i, const BlockChoice& c, const std::vector<int>& src_id) {
    std::vector<long long> k{(long long)i, c.d, c.s, c.dsi, c.dfm, -1};
    bool fc_dp = false, bin = false, sp = false;
    for (int a : c.active) {
        k.push_back(a);
        const int t = n.blocks[i].nodes[a].type;
        if (t == N_FC || t == N_DP) fc_dp = true;
        if (is_dense_binary(t)) bin = true;
        if (t == N_TRANS || t == N_EFC || t == N_DP) sp = true;
    }
    auto add = [&](const std::vector<int>& idx, long long tag, bool used) {
        k.push_back(-2 - tag);
        if (used) for (int j : idx) { k.push_back(j); k.push_back(src_id[j]); }
    };
    add(c.dense, 0, fc_dp); add(c.left, 1, bin); add(c.right, 2, bin); add(c.sparse, 3, sp);
    return k;
}

void forward_multi(Net& n, const int* choices, int n_cand, const float* int_x, const int64_t* cat_x, const float* emb_rows,
                   int B, float* logits, MultiStats& st) {
    n.tape_on = false;
    n.B = B;
    n.cat_x = cat_x;
    const int nd_ld = (n.nd + 3) & ~3;
    Var* x0 = n.var((int64_t)B * nd_ld, nd_ld != n.nd);
    if (nd_ld != n.nd) ck(nasrec_act_fwd(int_x, n.nd, B, n.nd, 0, x0->t, nd_ld, 0, n.st));
    else x0->t = const_cast<float*>(int_x);
    Var* sp0 = n.var((int64_t)B * n.F * E, emb_rows == nullptr);
    if (emb_rows) sp0->t = const_cast<float*>(emb_rows);
    else ck(nasrec_emb_gather_fwd(n.d_tables, n.d_rows, cat_x, sp0->t, B, n.F, n.d_err, n.st));
    struct Entry { DSrc d; SSrc s; };
    std::map<std::vector<long long>, int> ids;        // dependency description -> id of the computed block output
    std::vector<Entry> outs;
    const BlockDesc& last = n.blocks[n.num_blocks - 1];
    for (int c = 0; c < n_cand; ++c) {
        std::vector<BlockChoice> ch;
        for (int i = 0; i < n.num_blocks; ++i) ch.push_back(decode(choices + ((size_t)c * n.num_blocks + i) * CHOICE_STRIDE));
        for (auto& bc : ch) for (int a : bc.active) if (a < 0 || a >= 8) throw CallFailed(NASREC_EINVAL);
        std::vector<DSrc> dsrc{DSrc{x0, n.nd, nd_ld}};
        std::vector<SSrc> ssrc{SSrc{sp0, n.F, 0}};
        std::vector<bool> have{true};
        std::vector<int> src_id{-1};                  // the stem is the same tensor for every candidate
        const std::vector<bool> need = liveness(n, ch);
        for (int i = 0; i < n.num_blocks; ++i) {
            if (!need[i + 1]) {
                dsrc.push_back(DSrc{nullptr, 0, 0});
                ssrc.push_back(SSrc{nullptr, 0, 0});
                have.push_back(false);
                src_id.push_back(-3);
                continue;
            }
            const std::vector<long long> key = block_key(n, i, ch[i], src_id);
            auto it = ids.find(key);
            if (it == ids.end()) {
                DSrc d;
                SSrc s;
                run_block(n, i, ch[i], dsrc, ssrc, have, B, d, s);
                // the block's two output tensors are its first two arena allocations: everything behind them was scratch
                char* end = std::max((char*)(d.v->t + d.v->n), (char*)(s.v->t + s.v->n));
                n.act.off = (((size_t)(end - n.act.base)) + 255) & ~(size_t)255;
                outs.push_back(Entry{d, s});
                it = ids.emplace(key, (int)outs.size() - 1).first;
                ++st.computed;
            } else {
                ++st.reused;
            }
            dsrc.push_back(outs[it->second].d);
            ssrc.push_back(outs[it->second].s);
            have.push_back(true);
            src_id.push_back(it->second);
        }
        const size_t mark = n.act.off;
        const DSrc& dl = dsrc.back();
        const SSrc& sl = ssrc.back();
        const int64_t bs = (int64_t)(sl.s + sl.g) * E;
        Segs segs{Seg{dl.v, 0, dl.ld, dl.w, 0}, Seg{sl.v, 0, bs, (int64_t)sl.s * E, last.maxd}};
        if (sl.g) segs.push_back(Seg{sl.v, (int64_t)sl.s * E, bs, (int64_t)sl.g * E, last.maxd + (int64_t)last.maxs * E});
        LinArgs a; a.W = n.final_w; a.b = n.final_b; a.d_out = 1;
        Var* out = linear_ln(n, segs, B, a);
        ck((int)cudaMemcpyAsync(logits + (size_t)c * B, out->t, (size_t)B * 4, cudaMemcpyDeviceToDevice, n.st));
        n.act.off = mark;
    }
}

template <class F>
int guarded(F&& f) {
    int rc = 0;
    try { f(); }
    catch (const OutOfArena&) { rc = NASREC_ENOSPACE; }
    catch (const CallFailed& e) { rc = e.rc; }
    catch (const std::bad_alloc&) { rc = NASREC_ETOOBIG; }
    // the plane announcement is a hint for the calls THIS executor makes: never leave it behind for another caller of the
    // GEMM entry points (whose copy of the weight may have moved on without the planes)
    nasrec_set_weight_planes(nullptr, nullptr, nullptr, 0, 0, 0, 0);
    return rc;
}

}  // namespace

extern "C" {

// desc_i: [num_blocks, nd, F, final_w, final_b, F x emb param index, then per block:
//          num_nodes, maxd, maxs, dp_P, merger[4], fm[4], then num_nodes x (type, p[24])]
// params_*: per parameter weight pointer, optimizer-state pointer, numel, rows, cols, requires-grad flag.
void* nasrec_net_create(const int* desc_i, int desc_len, int n_params, float* const* w, float* const* state,
                        const int64_t* numel, const int* rows, const int* cols, const int* req,
                        const float* const* d_tables, const int64_t* d_rows, float* const* d_tables_rw,
                        float* const* d_states, int* d_err) {
    Net* n = new (std::nothrow) Net();
    if (!n || !desc_i || desc_len < 5) { delete n; return nullptr; }
    int o = 0;
    n->num_blocks = desc_i[o++]; n->nd = desc_i[o++]; n->F = desc_i[o++]; n->final_w = desc_i[o++]; n->final_b = desc_i[o++];
    for (int f = 0; f < n->F; ++f) n->emb_par.push_back(desc_i[o++]);
    for (int b = 0; b < n->num_blocks; ++b) {
        BlockDesc bd{};
        bd.num_nodes = desc_i[o++]; bd.maxd = desc_i[o++]; bd.maxs = desc_i[o++]; bd.dp_P = desc_i[o++];
        for (int k = 0; k < 4; ++k) bd.merger[k] = desc_i[o++];
        for (int k = 0; k < 4; ++k) bd.fm[k] = desc_i[o++];
        if (bd.num_nodes < 0 || bd.num_nodes > 8) { delete n; return nullptr; }
        for (int k = 0; k < bd.num_nodes; ++k) {
            bd.nodes[k].type = desc_i[o++];
            for (int q = 0; q < 24; ++q) bd.nodes[k].p[q] = desc_i[o++];
        }
        n->blocks.push_back(bd);
    }
    if (o != desc_len) { delete n; return nullptr; }
    for (int i = 0; i < n_params; ++i) { Par p{}; p.w = w[i]; p.st = state ? state[i] : nullptr; p.g = nullptr; p.n = numel[i]; p.rows = rows[i]; p.cols = cols[i]; p.req = req[i] != 0; n->par.push_back(p); }
    n->d_tables = d_tables; n->d_rows = d_rows; n->d_table_ptrs_rw = d_tables_rw; n->d_state_ptrs = d_states; n->d_err = d_err;
    return n;
}

void nasrec_net_destroy(void* net) {
    Net* n = (Net*)net;
    if (n && n->join_ev) cudaEventDestroy(n->join_ev);
    delete n;
}

int nasrec_net_set_arenas(void* net, void* act, int64_t act_bytes, void* pgrad, int64_t pgrad_bytes) {
    Net* n = (Net*)net;
    CHECK_ARG(n && act && pgrad && act_bytes > 0 && pgrad_bytes > 0);
    n->act = Arena{(char*)act, (size_t)act_bytes, 0, 0};
    n->pg = Arena{(char*)pgrad, (size_t)pgrad_bytes, 0, 0};
    n->pg_dirty = (size_t)pgrad_bytes;       // unknown contents: clear everything once
    // whatever an earlier step left in the old arenas is gone: make reduce / apply refuse instead of reading freed memory
    for (int pi : n->touched) n->par[pi].g = nullptr;
    n->touched.clear();
    n->emb_gout = nullptr;
    n->have_sparse = false;
    n->step_valid = false;
    n->res_rows = 0;
    n->res_partial = nullptr;
    return 0;
}

int nasrec_net_set_requires_grad(void* net, const int* req, int n_params) {
    Net* n = (Net*)net;
    CHECK_ARG(n && req && n_params == (int)n->par.size());
    for (int i = 0; i < n_params; ++i) n->par[i].req = req[i] != 0;
    return 0;
}

int nasrec_net_set_planes(void* net, float* const* hi, float* const* lo, const int64_t* ldp, const int* first, int n_params) {
    Net* n = (Net*)net;
    CHECK_ARG(n && n_params == (int)n->par.size());
    for (int i = 0; i < n_params; ++i) {
        Par& p = n->par[i];
        if (hi && lo && ldp && first && hi[i] && lo[i]) {
            CHECK_ARG(ldp[i] >= p.cols && (ldp[i] & 3) == 0 && first[i] >= 0 && first[i] <= p.cols);
            p.hi = hi[i]; p.lo = lo[i]; p.ldp = ldp[i]; p.first = first[i];
        } else {
            p.hi = p.lo = nullptr; p.ldp = 0; p.first = 0;
        }
    }
    return 0;
}

int nasrec_net_set_overlap(void* net, int on) {
    ((Net*)net)->overlap = on != 0;
    ((Net*)net)->late_join = on == 2;       // 2: the caller always runs nasrec_net_apply next and reads no gradient before it
    return 0;
}
int nasrec_net_set_defer_wgrad(void* net, int on) { ((Net*)net)->defer_wgrad = on != 0; return 0; }

int nasrec_net_set_reserve(void* net, int rows) {
    CHECK_ARG(net && rows >= 0);
    ((Net*)net)->reserve_rows = rows;
    return 0;
}

int nasrec_net_set_seal_callback(void* net, nasrec_seal_cb_t cb) {
    CHECK_ARG(net);
    ((Net*)net)->seal_cb = cb;
    return 0;
}

// Forward only (candidate scoring / inference).  emb_rows: optional pre-gathered [B,F,16] rows.
int nasrec_net_forward(void* net, const int* choice, const float* int_x, const int64_t* cat_x, const float* emb_rows,
                       int B, float* logits, void* stream) {
    Net* n = (Net*)net;
    CHECK_ARG(n && choice && int_x && (cat_x || emb_rows) && logits && B > 0);
    return guarded([&] {
        reset_step(*n, as_stream(stream));
        Var* out = forward(*n, choice, int_x, cat_x, emb_rows, B, false);
        ck((int)cudaMemcpyAsync(logits, out->t, (size_t)B * 4, cudaMemcpyDeviceToDevice, n->st));
    });
}

// Logits [n_cand, B] of n_cand subnets on ONE batch against the resident weights; blocks shared between candidates are
// computed once (see forward_multi).  stats (optional, host): {blocks computed, blocks reused}.
int nasrec_multi_subnet_eval(void* net, const int* choices, int n_cand, const float* int_x, const int64_t* cat_x,
                             const float* emb_rows, int B, float* logits, int* stats, void* stream) {
    Net* n = (Net*)net;
    CHECK_ARG(n && choices && n_cand > 0 && int_x && (cat_x || emb_rows) && logits && B > 0);
    MultiStats ms;
    const int rc = guarded([&] {
        reset_step(*n, as_stream(stream));
        n->step_valid = false;
        forward_multi(*n, choices, n_cand, int_x, cat_x, emb_rows, B, logits, ms);
    });
    if (stats) { stats[0] = ms.computed; stats[1] = ms.reused; }
    return rc;
}

// Forward + BCE + backward.  Afterwards the parameter gradients sit back to back in the pgrad arena
// (nasrec_net_grad_bucket) and the embedding gradient is available raw (nasrec_net_sparse_raw).
int nasrec_net_forward_backward(void* net, const int* choice, const float* int_x, const int64_t* cat_x, const float* y,
                                int B, float grad_scale, float* logits, float* loss, void* stream) {
    Net* n = (Net*)net;
    CHECK_ARG(n && choice && int_x && cat_x && y && logits && loss && B > 0);
    const int rc = guarded([&] {
        cudaStream_t st = as_stream(stream);
        reset_step(*n, st);
        n->step_valid = false;
        {   // reserve what sparse_reduce / apply will need, before any work is queued
            const int rows = n->reserve_rows > B ? n->reserve_rows : B;
            const int F = n->F;
            n->res_uniq = (int64_t*)n->act.alloc_bytes((size_t)F * rows * 8);
            n->res_nuniq = (int*)n->act.alloc_bytes((size_t)F * 4);
            n->res_row_grad = n->act.alloc((int64_t)F * rows * E);
            n->res_sumsq = n->act.alloc(F);
            n->res_scratch = (int*)n->act.alloc_bytes((size_t)F * (rows + 1) * 4);
            n->res_big_ws = nullptr;
            n->res_big_bytes = 0;
            if (rows > BIG_REDUCE_ROWS && F <= 31) {
                n->res_big_bytes = nasrec_emb_grad_sort_reduce_big_ws_bytes(rows, F);
                n->res_big_ws = n->act.alloc_bytes((size_t)n->res_big_bytes);
            }
            n->res_rows = rows;
            int64_t chunks = 0;
            for (const Par& p : n->par) chunks += (p.n + 16383) / 16384;      // >= nasrec_sumsq_ws_floats of any subset
            n->res_partial = n->act.alloc(chunks);
            n->res_partial_n = chunks;
        }
        if (n->join_pending) {              // a step that never reached nasrec_net_apply: its side-stream work ends before this one starts
            n->join_pending = false;
            ck((int)cudaStreamWaitEvent(st, n->join_ev, 0), 0);
        }
        if (n->pg_dirty) ck((int)cudaMemsetAsync(n->pg.base, 0, n->pg_dirty, st));
        Var* out = forward(*n, choice, int_x, cat_x, nullptr, B, true);
        out->g = n->act.alloc(B);
        ck(nasrec_bce_fwd_bwd(out->t, y, B, grad_scale, loss, out->g, st));
        ck((int)cudaMemcpyAsync(logits, out->t, (size_t)B * 4, cudaMemcpyDeviceToDevice, st));
        // Backward.  A parameter belongs to exactly one block (or the head), so when the tape index drops below a
        // block's first record every gradient written so far is final: seal that part of the bucket.
        size_t sealed = 0;
        int mark = (int)n->block_mark.size() - 1;
        nasrec_wgrad_defer(n->defer_wgrad ? 1 : 0);
        if (n->defer_wgrad) {
            // first-stage partials of every LayerNorm backward of this pass (<= 32 recorded per flush, <= 296 x 2 x 1024
            // floats each) stay in the arena until the batched final reduction of their block has been launched
            const long long lnd = (long long)std::min(296, (B + 3) / 4) * 2 * 1024 * 32;
            nasrec_internal_ln_defer_scratch(n->act.alloc(lnd), lnd);
        }
        // queued weight gradients run as one batched launch per block, on the side stream when one is attached: the batch
        // then overlaps the dY -> dX chain of the next block instead of extending the step by its ~90 us
        auto flush_wgrads = [&]() {
            if (nasrec_wgrad_pending() == 0) return;
            ck(nasrec_wgrad_flush(n->overlap ? nasrec_internal_fork_side(st) : st), 0);
        };
        for (size_t i = n->tape.size(); i-- > 0;) {
            n->tape[i]();
            if (mark >= 0 && i == n->block_mark[mark]) {
                --mark;
                flush_wgrads();
                if (n->seal_cb && n->pg.off >= sealed + n->seal_min_bytes) {       // few, large exchanges: each one costs host time
                    if (n->overlap) ck(nasrec_side_join(st), 0);     // the sealed range includes side-stream work
                    n->seal_cb((int64_t)sealed, (int64_t)(n->pg.off - sealed));
                    sealed = n->pg.off;
                }
            }
        }
        n->tape.clear();
        flush_wgrads();
        nasrec_wgrad_defer(0);
        nasrec_internal_ln_defer_scratch(nullptr, 0);
        // The last batch of parameter gradients (side stream) is needed by the optimizer only: without a seal callback the
        // join moves to nasrec_net_apply, so that the sorted-row reduction of the embedding gradient (nasrec_net_sparse_reduce,
        // 26 CTAs on the main stream) overlaps it (only when the caller asked for it: nasrec_net_set_overlap(net, 2)).
        n->join_pending = false;
        if (n->overlap && n->late_join && !n->seal_cb) {
            if (cudaStream_t side = nasrec_internal_side_stream()) {     // still attached here; the caller detaches it on return
                if (!n->join_ev) ck((int)cudaEventCreateWithFlags(&n->join_ev, cudaEventDisableTiming), 0);
                ck((int)cudaEventRecord(n->join_ev, side), 0);
                n->join_pending = true;
            }
        }
        if (n->overlap && !n->join_pending) ck(nasrec_side_join(st), 0);
        if (n->seal_cb && n->pg.off > sealed) n->seal_cb((int64_t)sealed, (int64_t)(n->pg.off - sealed));
        n->pg_dirty = n->pg.off;
        n->step_valid = true;
    });
    if (rc) {
        n->pg_dirty = n->pg.cap;       // a failed step may have scribbled anywhere in the bucket: clear all of it next time
        nasrec_wgrad_defer(0);         // ... and nothing of it may stay queued (turning deferral off drops the queue)
        nasrec_internal_ln_defer_scratch(nullptr, 0);
    }
    return rc;
}

int nasrec_net_grad_bucket(void* net, float** ptr, int64_t* nfloats) {
    Net* n = (Net*)net;
    CHECK_ARG(n && ptr && nfloats);
    *ptr = (float*)n->pg.base;
    *nfloats = (int64_t)(n->pg.off / 4);
    return 0;
}

int nasrec_net_sparse_raw(void* net, const int64_t** cat_x, float** gout) {
    Net* n = (Net*)net;
    CHECK_ARG(n && cat_x && gout);
    *cat_x = n->cat_x;
    *gout = n->emb_gout;
    return 0;
}

// Deterministic sorted-row reduction of the embedding gradient of B_all samples (the local batch, or the
// all-gathered global batch under data parallelism).  Pass NULLs to reduce the step's own raw gradient.
int nasrec_net_sparse_reduce(void* net, const int64_t* cat_all, const float* gout_all, int B_all, void* stream) {
    Net* n = (Net*)net;
    CHECK_ARG(n);
    return guarded([&] {
        const int64_t* cat = cat_all ? cat_all : n->cat_x;
        const float* go = gout_all ? gout_all : n->emb_gout;
        const int Bs = cat_all ? B_all : n->B;
        if (!n->step_valid) throw CallFailed(NASREC_EINVAL);      // no step, or its arenas were replaced
        if (!go) { n->have_sparse = false; return; }
        const int F = n->F;
        int* scratch;
        if (Bs <= n->res_rows) {
            n->uniq = n->res_uniq; n->nuniq = n->res_nuniq; n->row_grad = n->res_row_grad; n->sumsq = n->res_sumsq;
            scratch = n->res_scratch;
        } else {                                                  // the caller did not announce this many rows (nasrec_net_set_reserve)
            n->uniq = (int64_t*)n->act.alloc_bytes((size_t)F * Bs * 8);
            n->nuniq = (int*)n->act.alloc_bytes((size_t)F * 4);
            n->row_grad = n->act.alloc((int64_t)F * Bs * E);
            n->sumsq = n->act.alloc(F);
            scratch = (int*)n->act.alloc_bytes((size_t)F * (Bs + 1) * 4);
        }
        if (Bs > BIG_REDUCE_ROWS && F <= 31) {
            // large (all-gathered / KDD) batches: multi-CTA radix-sort reduction
            void* ws = n->res_big_ws;
            int64_t wb = n->res_big_bytes;
            const int64_t need = nasrec_emb_grad_sort_reduce_big_ws_bytes(Bs, F);
            if (!ws || wb < need || Bs > n->res_rows) { ws = n->act.alloc_bytes((size_t)need); wb = need; }
            ck(nasrec_emb_grad_sort_reduce_big(cat, n->d_rows, n->d_err, go, Bs, F, n->uniq, n->nuniq, n->row_grad, n->sumsq, ws, wb,
                                               as_stream(stream)), 10);
        } else {
            ck(nasrec_emb_grad_sort_reduce_checked(cat, n->d_rows, n->d_err, go, Bs, F, n->uniq, n->nuniq, n->row_grad, n->sumsq,
                                                   scratch, as_stream(stream)));
        }
        n->sB = Bs;
        n->have_sparse = true;
    });
}

// clip_grad_norm_(max_norm) over everything that received a gradient + Adagrad (train_utils.py:285-286).
// max_norm <= 0 disables clipping.  norm_out: 2 device floats (total norm, clip coefficient).
int nasrec_net_apply(void* net, float lr, float eps, float max_norm, float* norm_out, void* stream) {
    Net* n = (Net*)net;
    CHECK_ARG(n && norm_out);
    return guarded([&] {
        cudaStream_t st = as_stream(stream);
        if (!n->step_valid) throw CallFailed(NASREC_EINVAL);
        if (n->join_pending) {
            n->join_pending = false;
            ck((int)cudaStreamWaitEvent(st, n->join_ev, 0), 0);
        }
        std::vector<int> dense;
        for (int pi : n->ref_order) {
            bool is_emb = false;
            for (int e : n->emb_par) if (e == pi) is_emb = true;
            if (!is_emb && n->par[pi].g) dense.push_back(pi);
        }
        std::vector<float*> w, s, hi, lo;
        std::vector<const float*> g;
        std::vector<int64_t> sz, ldp;
        std::vector<int> cols, first;
        for (int pi : dense) {
            Par& p = n->par[pi];
            if (!p.st) throw CallFailed(NASREC_EINVAL);
            w.push_back(p.w); s.push_back(p.st); g.push_back(p.g); sz.push_back(p.n);
            hi.push_back(p.hi); lo.push_back(p.lo); ldp.push_back(p.ldp); cols.push_back(p.cols); first.push_back(p.first);
        }
        const float* coef = nullptr;
        if (max_norm > 0.f) {
            const int64_t need = nasrec_sumsq_ws_floats(sz.data(), (int)sz.size());
            float* partial = (n->res_partial && need <= n->res_partial_n) ? n->res_partial : n->act.alloc(need);
            ck(nasrec_grad_norm_clip(g.data(), sz.data(), (int)sz.size(), n->have_sparse ? n->sumsq : nullptr,
                                     n->have_sparse ? n->F : 0, max_norm, partial, norm_out, st), 2);
            coef = norm_out + 1;
        }
        if (!dense.empty())
            ck(nasrec_adagrad_multi_planes(w.data(), g.data(), s.data(), sz.data(), (int)sz.size(), lr, eps, coef, hi.data(),
                                           lo.data(), cols.data(), first.data(), ldp.data(), st));
        if (n->have_sparse)
            ck(nasrec_emb_rowwise_adagrad(n->uniq, n->nuniq, n->row_grad, n->d_table_ptrs_rw, n->d_state_ptrs, n->sB, n->F, lr,
                                          eps, coef, st));
    });
}

int64_t nasrec_net_launches(void) { return (int64_t)g_launches; }

int64_t nasrec_net_arena_high_water(void* net, int which) {
    Net* n = (Net*)net;
    if (!n) return -1;
    return (int64_t)(which == 0 ? n->act.high : n->pg.high);
}

}  // extern "C"

// Op-level entry points: one C call per operator direction instead of one per kernel.
// They only sequence the kernel launchers of gemm.cu / ln.cu on the caller's stream; the point is
// host-side cost: at B=512 a supernet step is ~150 launches and the Python->C boundary dominated it.
#include "common.cuh"

namespace {
struct Targets {
    nasrec_seg_t seg[NASREC_MAX_SEGS];
    int flag[NASREC_MAX_SEGS];
    int n = 0;
};
// grad targets: ptr == null -> no gradient wanted; flag 0 -> overwrite, 1 -> accumulate.  Fresh and accumulated targets go
// into ONE dgrad launch (per-problem accumulate, nasrec_internal_set_dgrad_flags): a launch costs ~7 us whatever its size.
void wanted_targets(const nasrec_seg_t* dsegs, const int* acc_flags, int nseg, Targets& t) {
    for (int i = 0; i < nseg; ++i) {
        if (!dsegs[i].ptr || dsegs[i].width == 0) continue;
        t.seg[t.n] = dsegs[i];
        t.flag[t.n] = acc_flags ? (acc_flags[i] != 0) : 0;
        ++t.n;
    }
}

// ---- fork / join between the caller's stream and the optional side stream ----------------------
// Weight gradients are consumed only by the optimizer at the end of the step, so they need not sit on
// the critical path dY -> dX of the backward pass.  With a side stream attached, each backward entry
// point forks after its LayerNorm-backward kernel: wgrad (+ bias grad) go to the side stream, dgrad
// stays on the caller's stream; nasrec_side_join() makes the caller's stream wait for everything forked.
// Works under CUDA-graph capture (the side stream joins the capture through the event wait).
constexpr int NEV = 64;
cudaEvent_t g_ev[NEV];
int g_ev_made = 0, g_ev_next = 0;
bool g_forked = false;

cudaEvent_t next_event() {
    if (g_ev_made < NEV) {
        cudaEventCreateWithFlags(&g_ev[g_ev_made], cudaEventDisableTiming);
        return g_ev[g_ev_made++];
    }
    g_ev_next = (g_ev_next + 1) % NEV;
    return g_ev[g_ev_next];
}

// returns the stream wgrad work should use
cudaStream_t fork_for_wgrad(cudaStream_t main) {
    cudaStream_t side = nasrec_internal_side_stream();
    if (!side || side == main) return main;
    cudaEvent_t e = next_event();
    cudaEventRecord(e, main);
    cudaStreamWaitEvent(side, e, 0);
    g_forked = true;
    return side;
}
}  // namespace

cudaStream_t nasrec_internal_fork_side(cudaStream_t main) { return fork_for_wgrad(main); }

extern "C" {

int nasrec_set_side_stream(void* stream) {
    nasrec_internal_set_side_stream((cudaStream_t)stream);
    g_forked = false;
    return 0;
}

int nasrec_side_join(void* stream) {
    cudaStream_t side = nasrec_internal_side_stream();
    if (!side || !g_forked) return 0;
    cudaEvent_t e = next_event();
    cudaEventRecord(e, side);
    cudaStreamWaitEvent(as_stream(stream), e, 0);
    g_forked = false;
    return nasrec_launch_status();
}

// y[:, :d_out] (+)= act(LN(concat(segs) @ W[n_off:n_off+N].T + bias))   (LN skipped when gamma == null)
int nasrec_linear_ln_fwd(const nasrec_seg_t* segs, int nseg, const float* W, int64_t ldw, int n_off, int N,
                         const float* bias, const float* gamma, const float* beta, float eps, int relu, int d_out,
                         float* z, float* y, int64_t ldy, float* mean, float* rstd, int accumulate, int M,
                         void* stream) {
    int rc = nasrec_seg_linear_fwd(segs, nseg, W, ldw, n_off, N, bias, z, N, M, stream);
    if (rc) return rc;
    if (gamma) return nasrec_ln_fwd(z, N, M, N, gamma, beta, eps, relu, d_out, y, ldy, mean, rstd, accumulate, stream);
    return nasrec_act_fwd(z, N, M, N, relu, y, ldy, accumulate, stream);
}

// Backward of the above.  dz: [M,N] scratch.  dW/dbias/dgamma/dbeta/dsegs[i].ptr may be null (not wanted).
// dsegs must not contain the same target twice, segs must not meet the same W columns twice.
int nasrec_linear_ln_bwd(const float* dy, int64_t lddy, int d_out, const float* z, int M, int N, const float* gamma,
                         const float* beta, const float* mean, const float* rstd, int relu, const nasrec_seg_t* segs,
                         const nasrec_seg_t* dsegs, const int* dseg_accumulate, int nseg, const float* W, int64_t ldw,
                         int n_off, float* dW, float* dbias, float* dgamma, float* dbeta, float* dz, void* stream) {
    int rc;
    if (gamma) rc = nasrec_ln_bwd(dy, lddy, d_out, z, N, M, N, gamma, beta, mean, rstd, relu, dz, N, dgamma, dbeta, 0, stream);
    else rc = nasrec_act_bwd(dy, lddy, z, N, M, N, relu, dz, N, stream);
    if (rc) return rc;
    if (dW || dbias) {
        void* ws_stream = fork_for_wgrad(as_stream(stream));
        if (dW) {
            rc = nasrec_seg_linear_wgrad(dz, N, N, segs, nseg, dW, ldw, n_off, M, 0, ws_stream);
            if (rc) return rc;
        }
        if (dbias) {
            rc = nasrec_colsum(dz, N, M, N, dbias + n_off, 0, ws_stream);
            if (rc) return rc;
        }
    }
    Targets t;
    wanted_targets(dsegs, dseg_accumulate, nseg, t);
    if (t.n) {
        nasrec_internal_set_dgrad_flags(t.flag);
        rc = nasrec_seg_linear_dgrad(dz, N, N, W, ldw, n_off, t.seg, t.n, M, 0, stream);
        nasrec_internal_set_dgrad_flags(nullptr);
        if (rc) return rc;
    }
    return 0;
}

// y[b, p<p_out, :] (+)= act(LN_P(W @ concat_rows(segs)[b] + bias))
int nasrec_sproj_ln_fwd(const nasrec_seg_t* segs, int nseg, const float* W, int64_t ldw, int P, const float* bias,
                        const float* gamma, const float* beta, float eps, int relu, int p_out, float* z, float* y,
                        int64_t y_bstride, float* mean, float* rstd, int accumulate, int B, void* stream) {
    int rc = nasrec_sproj_fwd(segs, nseg, W, ldw, P, bias, z, (int64_t)P * NASREC_EMB_DIM, B, stream);
    if (rc) return rc;
    if (gamma)
        return nasrec_ln3_fwd(z, (int64_t)P * NASREC_EMB_DIM, B, P, gamma, beta, eps, relu, p_out, y, y_bstride, mean,
                              rstd, accumulate, stream);
    return nasrec_act_fwd(z, (int64_t)P * NASREC_EMB_DIM, B, p_out * NASREC_EMB_DIM, relu, y, y_bstride, accumulate, stream);
}

int nasrec_sproj_ln_bwd(const float* dy, int64_t dy_bstride, int p_out, const float* z, int B, int P, const float* gamma,
                        const float* beta, const float* mean, const float* rstd, int relu, const nasrec_seg_t* segs,
                        const nasrec_seg_t* dsegs, const int* dseg_accumulate, int nseg, const float* W, int64_t ldw,
                        float* dW, float* dbias, float* dgamma, float* dbeta, float* dz, float* ws, void* stream) {
    const int64_t zbs = (int64_t)P * NASREC_EMB_DIM;
    int rc;
    if (gamma) rc = nasrec_ln3_bwd(dy, dy_bstride, p_out, z, zbs, B, P, gamma, beta, mean, rstd, relu, dz, zbs, dgamma, dbeta, 0, stream);
    else rc = nasrec_act_bwd(dy, dy_bstride, z, zbs, B, P * NASREC_EMB_DIM, relu, dz, zbs, stream);
    if (rc) return rc;
    if (dW || dbias) {
        void* ws_stream = fork_for_wgrad(as_stream(stream));
        if (dW) {
            rc = nasrec_sproj_wgrad(dz, zbs, P, segs, nseg, dW, ldw, B, 0, ws, ws_stream);
            if (rc) return rc;
        }
        if (dbias) {
            rc = nasrec_sproj_bias_grad(dz, zbs, P, B, dbias, 0, ws_stream);
            if (rc) return rc;
        }
    }
    Targets t;
    wanted_targets(dsegs, dseg_accumulate, nseg, t);
    if (t.n) {
        nasrec_internal_set_dgrad_flags(t.flag);
        rc = nasrec_sproj_dgrad(dz, zbs, P, W, ldw, t.seg, t.n, B, 0, stream);
        nasrec_internal_set_dgrad_flags(nullptr);
        if (rc) return rc;
    }
    return 0;
}

}  // extern "C"

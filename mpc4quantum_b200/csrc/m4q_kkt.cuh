// Pivoted stage-wise KKT solve: the QP's last resort when the Riccati recursion cannot certify.
//
// Why it exists (profiles/r2_h100_order1_analysis.md): with the order-1 (Euler) model every
// transition frequency gives |eig A_t| > 1, at H = 100 the open-loop transition has norm 4e13 and neither the
// cost-to-go (entries ~1e27) nor any rollout through the dynamics keeps digits.  The reference's solver (cvxpy -> OSQP,
// optimize.py:59; our CPU oracle: a sparse LU of the KKT matrix) never propagates through prod A_t, because states AND
// costates are unknowns of one banded system.  This is the device equivalent: the KKT system of the working set,
//   stationarity  2 R_t (u_t - ub_t) + B_t^T lam_{t+1} = 0     (free controls; pinned ones: u_t[i] = bound)
//   dynamics      x_{t+1} - A_t x_t - B_t u_t = D_t
//   costate       lam_t - 2 Q_t x_t - A_t^T lam_{t+1} = -2 Q_t r_t        (lam_H = 2 Qf (x_H - r_H))
// ordered stage by stage -- block s holds [u_{s-1} | x_s | lam_s], W = 2N + M unknowns -- is almost block diagonal and
// is eliminated block column by block column with ROW PARTIAL PIVOTING over every row that touches the column: the N
// rows carried over from the previous block, the block's own stationarity and costate rows and the next stage's
// dynamics rows (3N + M candidates, two block columns + the right-hand side wide).  Pivot rows go to a per-warp array
// in global memory (H W (2W + 2) doubles: 2.4 MB for the transmon at H = 100), back-substitution runs block by block
// from the end, and the multipliers of the box constraints come from the solve's own costates, so they carry the
// accuracy of the solve (1e-7 at worst on the captured H = 100 problems, tools/analysis/abd_full.py) instead of the
// noise of an adjoint sweep through the unstable dynamics.
//
// Active set: primal-dual rounds from an EMPTY working set, the iteration the oracle uses (oracle/restate.py:
// _active_set); on the captured problems it settles in 1..25 rounds where the warm-started rounds cycle
// (tools/analysis/abd_active.py).  Weakly active bounds get the same hysteresis as in qp_solve.
//
// One warp per member as everywhere else; the elimination window (3N + M rows) lives in the same per-warp global
// array and is served by L1/L2.  This path is cold for every configuration that the Riccati path certifies.
#pragma once

namespace m4q {

template <class CF> struct Kkt {
    static constexpr int N = CF::N, M = CF::M, C = CF::C;
    static constexpr int W = 2 * N + M;        // unknowns per block: u_{s-1} | x_s | lam_s
    static constexpr int ROWS = W + N;         // candidate rows of a block column
    static constexpr int RHS = 2 * W;          // column of the right-hand side
    static constexpr int LD = 2 * W + 2;       // row length (even)
    static constexpr int CH = cdiv(LD, 32);    // column chunks of a row per lane
    static constexpr int CR = cdiv(ROWS, 32);  // rows of a column per lane
    static constexpr int CW = cdiv(W, 32);     // unknowns of a block per lane
    __host__ __device__ static constexpr long long doubles(int H) { return (long long)(ROWS + (long long)H * W) * LD; }
};

// realified entry (k, j) of a complex C x C block stored with row stride CA: [[Re, -Im], [Im, Re]]
template <class CF> __device__ __forceinline__ double realified(const double2 *At, int k, int j) {
    constexpr int C = CF::C;
    const double2 a = At[(k % C) * Rec<CF>::CA + (j % C)];
    return ((k < C) == (j < C)) ? a.x : (k >= C ? a.y : -a.y);
}

// One equality-constrained solve for the working set in slab.mask.  Out: Uo (slab), Xo (workspace), the gradient of the
// objective w.r.t. every control in slab.kk (= the multipliers on pinned controls, ~0 on free ones).
// Returns false if a pivot vanished or the solution is not finite.
template <class CF, bool FUSED>
__device__ __noinline__ bool kkt_solve(SlabRef sr, const QPData &qp_in, double *kkt, int lane) {
    using K_ = Kkt<CF>;
    using R_ = Rec<CF>;
    constexpr int N = CF::N, M = CF::M, W = K_::W, LD = K_::LD, CH = K_::CH, CR = K_::CR, CW = K_::CW, RHS = K_::RHS;
    const Slab<CF> s = slab_view<CF>(sr);
    const QPData qp = localize<FUSED>(qp_in);
    const int H = sr.H;
    double *win = kkt;                         // [ROWS][LD] elimination window
    double *Ust = kkt + K_::ROWS * LD;         // [H][W][LD] pivot rows
    bool ok = true;

    // rows of block s (1-based): fills window rows row0.. with stat_s, cos_s and (s < H) dyn_{s+1}; columns of block s
    // at 0, of block s+1 at W
    auto fill_row = [&](int row, int kind, int s_, int k) {
        // kind 0: stat (control k of stage s_-1), 1: cos (state k of stage s_), 2: dyn_{s_+1} (state k), 3: dyn_1
        double *wr = win + row * LD;
        const int t = s_ - 1;
        const double *rec_t = ws_rec<CF>(sr, t < 0 ? 0 : t);
        const double *rec_s = ws_rec<CF>(sr, s_ < H ? s_ : H - 1);
#pragma unroll
        for (int cc = 0; cc < CH; ++cc) {
            const int c = lane + 32 * cc;
            if (c >= LD) continue;
            double v = 0.0;
            if (kind == 0) {
                const int mk = s.mask[t * M + k];
                if (mk) {
                    if (c == k) v = 1.0;
                    else if (c == RHS) v = mk == 1 ? box_lo(s, qp.sat, t, k) : box_hi(s, qp.sat, t, k);
                } else {
                    if (c < M) v = 2.0 * qp.R[t * qp.r_stride + k * M + c];
                    else if (c >= M + N && c < W) v = rec_t[R_::B + R_::pair(k, c - M - N)];
                    else if (c == RHS) v = 2.0 * qp.Rub[t * M + k];
                }
            } else if (kind == 1) {
                const double *Qs = s_ == H ? qp.Qf : qp.Q + s_ * qp.q_stride;
                if (c >= M && c < M + N) v = -2.0 * Qs[k * N + (c - M)];
                else if (c >= M + N && c < W) v = (c - M - N == k) ? 1.0 : 0.0;
                else if (c >= W + M + N && c < 2 * W) {
                    if (s_ < H) v = -realified<CF>(reinterpret_cast<const double2 *>(rec_s + R_::AT), c - W - M - N, k);
                } else if (c == RHS) v = -2.0 * (s_ == H ? qp.qlinf[k] : qp.qlin[s_ * N + k]);
            } else if (kind == 2) {
                if (c >= M && c < M + N) v = -realified<CF>(reinterpret_cast<const double2 *>(rec_s + R_::AT), k, c - M);
                else if (c >= W && c < W + M) v = -rec_s[R_::B + R_::pair(c - W, k)];
                else if (c == W + M + k) v = 1.0;
                else if (c == RHS) v = rec_s[R_::D + k];
            } else {
                if (c < M) v = -rec_t[R_::B + R_::pair(c, k)];
                else if (c == M + k) v = 1.0;
                else if (c == RHS) v = rec_t[R_::D + k];   // + (A_0 x_0)[k], added by the caller
            }
            wr[c] = v;
        }
    };

    // ---- block 1: the "carried" rows are the dynamics of stage 0 (x_0 is data)
    {
        const double2 *A0 = reinterpret_cast<const double2 *>(ws_rec<CF>(sr, 0) + R_::AT);
        double ax = 0.0;
        if (lane < N)
            for (int j = 0; j < N; ++j) ax = fma(realified<CF>(A0, lane, j), s.x0[j], ax);
#pragma unroll 1
        for (int k = 0; k < N; ++k) {
            fill_row(k, 3, 1, k);
            const double axk = __shfl_sync(FULL, ax, k);
            __syncwarp();
            if (lane == 0) win[k * LD + RHS] += axk;
        }
    }
    __syncwarp();

#pragma unroll 1
    for (int sblk = 1; sblk <= H; ++sblk) {
        const int nrows = sblk < H ? K_::ROWS : W;
#pragma unroll 1
        for (int i = 0; i < M; ++i) fill_row(N + i, 0, sblk, i);
#pragma unroll 1
        for (int k = 0; k < N; ++k) fill_row(N + M + k, 1, sblk, k);
        if (sblk < H) {
#pragma unroll 1
            for (int k = 0; k < N; ++k) fill_row(W + k, 2, sblk, k);
        }
        __syncwarp();
        double *Us = Ust + (size_t)(sblk - 1) * W * LD;
        // ---- eliminate the W columns of this block
#pragma unroll 1
        for (int k = 0; k < W; ++k) {
            double a[CR];
            double best = -1.0;
            int bi = 0;
#pragma unroll
            for (int q = 0; q < CR; ++q) {
                const int r = k + lane + 32 * q;
                a[q] = r < nrows ? win[r * LD + k] : 0.0;
                const double av = r < nrows ? fabs(a[q]) : -1.0;
                if (av > best) {
                    best = av;
                    bi = lane + 32 * q;
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ob = __shfl_xor_sync(FULL, best, o);
                const int oi = __shfl_xor_sync(FULL, bi, o);
                if (ob > best || (ob == best && oi < bi)) {
                    best = ob;
                    bi = oi;
                }
            }
            if (!(best > 0.0) || !isfinite(best)) ok = false;
            const int prow = k + bi;
            // pivot row -> registers and the pivot-row store; what sat in row k moves to the pivot's place
            double pr[CH];
#pragma unroll
            for (int cc = 0; cc < CH; ++cc) {
                const int c = lane + 32 * cc;
                pr[cc] = c < LD ? win[prow * LD + c] : 0.0;
            }
            if (bi != 0) {
#pragma unroll
                for (int cc = 0; cc < CH; ++cc) {
                    const int c = lane + 32 * cc;
                    if (c < LD) win[prow * LD + c] = win[k * LD + c];
                }
            }
#pragma unroll
            for (int cc = 0; cc < CH; ++cc) {
                const int c = lane + 32 * cc;
                if (c < LD) Us[k * LD + c] = pr[cc];
            }
            double akk = __shfl_sync(FULL, a[0], 0);
            double pv = 0.0;
#pragma unroll
            for (int q = 0; q < CR; ++q) {
                const double cand = __shfl_sync(FULL, a[q], bi & 31);
                if ((bi >> 5) == q) pv = cand;
            }
            if (bi != 0 && lane == (bi & 31)) {
#pragma unroll
                for (int q = 0; q < CR; ++q)
                    if ((bi >> 5) == q) a[q] = akk;
            }
            const double ipv = 1.0 / pv;
            __syncwarp();
            // rank-1 update of the rows below, RB rows per trip so that their loads are in flight together (the window
            // is served by L1 / L2: a row at a time is one memory round trip per row); rows whose entry in column k is
            // zero are skipped (warp-uniform test)
            constexpr int RB = 8;
#pragma unroll 1
            for (int j0 = 1; k + j0 < nrows; j0 += RB) {
                double lj[RB], v[RB][CH];
#pragma unroll
                for (int b = 0; b < RB; ++b) {
                    const int j = j0 + b;
                    double cand = 0.0;
#pragma unroll
                    for (int q = 0; q < CR; ++q) {
                        const double cq = __shfl_sync(FULL, a[q], j & 31);
                        if ((j >> 5) == q) cand = cq;
                    }
                    lj[b] = (k + j < nrows) ? cand * ipv : 0.0;
                }
#pragma unroll
                for (int b = 0; b < RB; ++b) {
                    const double *wr = win + (k + j0 + b) * LD;
#pragma unroll
                    for (int cc = 0; cc < CH; ++cc) {
                        const int c = lane + 32 * cc;
                        v[b][cc] = (lj[b] != 0.0 && c > k && c < LD - 1) ? wr[c] : 0.0;
                    }
                }
#pragma unroll
                for (int b = 0; b < RB; ++b) {
                    double *wr = win + (k + j0 + b) * LD;
#pragma unroll
                    for (int cc = 0; cc < CH; ++cc) {
                        const int c = lane + 32 * cc;
                        if (lj[b] != 0.0 && c > k && c < LD - 1) wr[c] = fma(-lj[b], pr[cc], v[b][cc]);
                    }
                }
            }
            __syncwarp();
        }
        // ---- carry: rows W.., columns of block s+1 and the right-hand side -> rows 0.., columns of the next block
        if (sblk < H) {
#pragma unroll 1
            for (int i = 0; i < N; ++i) {
                double v[CH];
#pragma unroll
                for (int cc = 0; cc < CH; ++cc) {
                    const int c = lane + 32 * cc;
                    v[cc] = 0.0;
                    if (c < W) v[cc] = win[(W + i) * LD + W + c];
                    else if (c == RHS) v[cc] = win[(W + i) * LD + RHS];
                }
#pragma unroll
                for (int cc = 0; cc < CH; ++cc) {
                    const int c = lane + 32 * cc;
                    if (c < LD) win[i * LD + c] = v[cc];
                }
            }
            __syncwarp();
        }
    }

    // ---- back-substitution, block by block from the end; z of the next block stays in registers
    double zn[CW];
#pragma unroll
    for (int q = 0; q < CW; ++q) zn[q] = 0.0;
    double *Xo = ws_Xo<CF>(sr);
    double *zb = s.scr;   // W doubles of scratch (the factor scratch is dead here)
    bool finite = true;
#pragma unroll 1
    for (int sblk = H; sblk >= 1; --sblk) {
        const double *Us = Ust + (size_t)(sblk - 1) * W * LD;
        double b[CW];
#pragma unroll
        for (int q = 0; q < CW; ++q) {
            const int r = lane + 32 * q;
            b[q] = r < W ? Us[r * LD + RHS] : 0.0;
        }
        if (sblk < H) {
#pragma unroll 1
            for (int c = 0; c < W; ++c) {
                double zc = 0.0;
#pragma unroll
                for (int q = 0; q < CW; ++q) {
                    const double cand = __shfl_sync(FULL, zn[q], c & 31);
                    if ((c >> 5) == q) zc = cand;
                }
#pragma unroll
                for (int q = 0; q < CW; ++q) {
                    const int r = lane + 32 * q;
                    if (r < W) b[q] = fma(-Us[r * LD + W + c], zc, b[q]);
                }
            }
        }
#pragma unroll 1
        for (int k = W - 1; k >= 0; --k) {
            double zk = 0.0;
#pragma unroll
            for (int q = 0; q < CW; ++q) {
                const int r = lane + 32 * q;
                double cand = (r == k) ? b[q] / Us[k * LD + k] : 0.0;
                if (r == k) b[q] = cand;
                cand = __shfl_sync(FULL, cand, k & 31);
                if ((k >> 5) == q) zk = cand;
            }
#pragma unroll
            for (int q = 0; q < CW; ++q) {
                const int r = lane + 32 * q;
                if (r < k) b[q] = fma(-Us[r * LD + k], zk, b[q]);
            }
        }
#pragma unroll
        for (int q = 0; q < CW; ++q) {
            zn[q] = b[q];
            const int r = lane + 32 * q;
            if (r < W) {
                zb[r] = b[q];
                finite &= isfinite(b[q]) != 0;
            }
        }
        __syncwarp();
        // unpack: controls, state, and the gradient 2 R (u - ub) + B^T lam of stage t = s - 1
        const int t = sblk - 1;
        const double *rec = ws_rec<CF>(sr, t);
        if (lane < N) Xo[sblk * N + lane] = zb[M + lane];
        double g[M];
#pragma unroll
        for (int i = 0; i < M; ++i) g[i] = lane < N ? rec[R_::B + R_::pair(i, lane)] * zb[M + N + lane] : 0.0;
        warp_sum_vec<M>(g, lane);
        if (lane == 0) {
            const double *Rt = qp.R + t * qp.r_stride;
#pragma unroll
            for (int i = 0; i < M; ++i) {
                const int mk = s.mask[t * M + i];
                const double u = mk == 1 ? box_lo(s, qp.sat, t, i) : (mk == 2 ? box_hi(s, qp.sat, t, i) : zb[i]);
                s.Uo[t * M + i] = u;
            }
#pragma unroll
            for (int i = 0; i < M; ++i) {
                double gi = g[i];
#pragma unroll
                for (int j = 0; j < M; ++j) gi = fma(2.0 * Rt[i * M + j], s.Uo[t * M + j] - qp.ub[t * M + j], gi);
                s.kk[t * M + i] = gi;
            }
        }
        __syncwarp();
    }
    if (lane < N) Xo[lane] = s.x0[lane];
    __syncwarp();
    return ok && !__any_sync(FULL, !finite);
}

// Active set on top of kkt_solve.  First primal-dual rounds from an empty working set (the oracle's iteration,
// oracle/restate.py: _active_set_kkt_multipliers); QPs whose unconstrained optimum lies ten box widths outside the
// box make those rounds erratic, so after 30 of them the textbook primal method takes over (Nocedal & Wright alg.
// 16.3 for a box: feasible iterates, the blocking bound of every step is added, all wrong-signed multipliers are
// dropped at a stationary point) -- it cannot cycle.  Returns 0 (KKT point found; z, y set for the next warm
// start), 2 (iterations exhausted) or 3 (singular / non-finite).
template <class CF, bool FUSED>
__device__ __noinline__ int kkt_active_set(SlabRef sr, const QPData &qp_in, const QPSet &set, int lane, Counters &cnt) {
    constexpr int M = CF::M;
    const Slab<CF> s = slab_view<CF>(sr);
    const QPData qp = localize<FUSED>(qp_in);
    const int HM = sr.H * M;
    // set.kkt is the array of all resident warps' workspaces; this warp's slice follows the launch geometry
    double *kkt = set.kkt + (size_t)(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * Kkt<CF>::doubles(sr.H);
    bool found = false;
#pragma unroll 1
    for (int e = lane; e < HM; e += 32) {
        s.mask[e] = 0;
        s.flips[e] = 0;
    }
    __syncwarp();
#pragma unroll 1
    for (int round = 0; round < 30 && !found; ++round) {
        cnt.kkt++;
        if (!kkt_solve<CF, FUSED>(sr, qp_in, kkt, lane)) return 3;
        double gmax = 0.0;
#pragma unroll 1
        for (int e = lane; e < HM; e += 32) gmax = fmax(gmax, fabs(s.kk[e]));
        const double gs = fmax(1.0, warp_max(gmax));
        bool changed = false;
#pragma unroll 1
        for (int e = lane; e < HM; e += 32) {
            const int t = e / M, i = e % M;
            const int mk = s.mask[e];
            const double u = s.Uo[e], g = s.kk[e];
            int nm = mk;
            if (mk == 0) {
                if (u < box_lo(s, qp.sat, t, i) - 1e-12) nm = 1;
                else if (u > box_hi(s, qp.sat, t, i) + 1e-12) nm = 2;
            } else {
                const double gn = mk == 1 ? -g : g;   // > 0: the multiplier has the wrong sign
                if (gn > 1e-10 * gs) {
                    const int fl = s.flips[e];
                    if (fl < 2 || gn > 1e-5 * gs) {
                        nm = 0;
                        s.flips[e] = fl + 1;
                    }
                }
            }
            if (nm != mk) {
                s.mask[e] = nm;
                changed = true;
            }
        }
        __syncwarp();
        found = !__any_sync(FULL, changed);
    }
    if (!found) {
        // ---- primal method from the feasible point z = clip(0)
#pragma unroll 1
        for (int e = lane; e < HM; e += 32) {
            const int t = e / M, i = e % M;
            const double lo = box_lo(s, qp.sat, t, i), hi = box_hi(s, qp.sat, t, i);
            const double z = fmin(fmax(0.0, lo), hi);
            s.z[e] = z;
            s.mask[e] = z <= lo ? 1 : (z >= hi ? 2 : 0);
        }
        __syncwarp();
#pragma unroll 1
        for (int it = 0; it < 40 * HM + 100 && !found; ++it) {
            cnt.kkt++;
            if (!kkt_solve<CF, FUSED>(sr, qp_in, kkt, lane)) return 3;
            double smax = 0.0, umax = 0.0, gmax = 0.0;
#pragma unroll 1
            for (int e = lane; e < HM; e += 32) {
                smax = fmax(smax, fabs(s.Uo[e] - s.z[e]));
                umax = fmax(umax, fabs(s.z[e]));
                gmax = fmax(gmax, fabs(s.kk[e]));
            }
            smax = warp_max(smax);
            umax = warp_max(umax);
            if (smax <= 1e-11 * fmax(1.0, umax)) {
                const double gs = fmax(1.0, warp_max(gmax));
                bool drop = false;
#pragma unroll 1
                for (int e = lane; e < HM; e += 32) {
                    const int mk = s.mask[e];
                    const double w = mk == 1 ? -s.kk[e] : (mk == 2 ? s.kk[e] : 0.0);
                    if (w > 1e-9 * gs) {
                        s.mask[e] = 0;
                        drop = true;
                    }
                }
                __syncwarp();
                found = !__any_sync(FULL, drop);
                continue;
            }
            // ratio test over the free controls
            double rmin = 1.0;
            int rk = -1;
#pragma unroll 1
            for (int e = lane; e < HM; e += 32) {
                if (s.mask[e]) continue;
                const int t = e / M, i = e % M;
                const double z = s.z[e], st = s.Uo[e] - z;
                double r = 2.0;
                if (st < 0.0) r = (box_lo(s, qp.sat, t, i) - z) / st;
                else if (st > 0.0) r = (box_hi(s, qp.sat, t, i) - z) / st;
                if (r < rmin) {
                    rmin = r;
                    rk = e;
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double orr = __shfl_xor_sync(FULL, rmin, o);
                const int ok = __shfl_xor_sync(FULL, rk, o);
                if (orr < rmin || (orr == rmin && ok >= 0 && (rk < 0 || ok < rk))) {
                    rmin = orr;
                    rk = ok;
                }
            }
            rmin = fmax(rmin, 0.0);
#pragma unroll 1
            for (int e = lane; e < HM; e += 32) {
                const int t = e / M, i = e % M;
                const double z = s.z[e], st = s.Uo[e] - z;
                if (rk < 0) s.z[e] = s.Uo[e];
                else if (e == rk) {
                    s.mask[e] = st < 0.0 ? 1 : 2;
                    s.z[e] = st < 0.0 ? box_lo(s, qp.sat, t, i) : box_hi(s, qp.sat, t, i);
                } else if (!s.mask[e]) s.z[e] = fma(rmin, st, z);
            }
            __syncwarp();
        }
    }
    if (!found) return 2;
    const double inv_rho = 1.0 / set.rho;
#pragma unroll 1
    for (int e = lane; e < HM; e += 32) {
        const int t = e / M, i = e % M;
        s.z[e] = fmin(fmax(s.Uo[e], box_lo(s, qp.sat, t, i)), box_hi(s, qp.sat, t, i));
        s.Uo[e] = s.z[e];
        s.y[e] = s.mask[e] ? -s.kk[e] * inv_rho : 0.0;
    }
    __syncwarp();
    return 0;
}

}   // namespace m4q

/*
 * pdeb_pair.cuh — "pair family": TWO adjacent lanes own one trajectory of an even mesh N.
 *
 * Lane A (even lane) holds rows 1..N/2 counted from the Dirichlet end, lane B rows N..N/2+1 counted
 * from the loaded free end, so both run the SAME instruction stream on mirrored local indices
 * l = 1..H (H = N/2): eliminate from the outer end towards the middle, solve the 2x2 system that
 * couples the two middle rows (one 64-bit shuffle), substitute back outwards.  Compared with the
 * register family (pdeb_core.h, one thread per trajectory, twisted factorisation) each thread carries
 * half the state and half the factors, which doubles the number of resident warps per SM; the
 * arithmetic per trajectory is the same.
 *
 * Mathematics as in pdeb_core.h: heat = generalised trapezoidal d-form on the deviation w = d - g0, wave =
 * Newmark a-form, STL clauses reduced on the fly by one masked max-reduction per clause
 * (reference call sites: heatlinfem.heatlinfem_mix / mechlinfem.mechlinfem / logic.APCont / cspec in
 * utils/few_shot_prompts/examples/one_d_heat/examples.jsonl:1-3, one_d_wave/examples.jsonl:1-3).
 */
#ifndef PDEB_PAIR_CUH
#define PDEB_PAIR_CUH

#include <cuda_runtime.h>

#include "pdeb_core.h"

#ifndef PDEB_PAIR_THREADS
#define PDEB_PAIR_THREADS 128
#endif
#ifndef PDEB_PAIR_MINBLOCKS
#define PDEB_PAIR_MINBLOCKS 4
#endif
#define PDEB_ALL 0xffffffffu

__device__ __forceinline__ double pdeb_partner(double x) { return __shfl_xor_sync(PDEB_ALL, x, 1); }
__device__ __forceinline__ float pdeb_partner(float x) { return __shfl_xor_sync(PDEB_ALL, x, 1); }

template <int KIND, int N, typename real>
__global__ void __launch_bounds__(PDEB_PAIR_THREADS, PDEB_PAIR_MINBLOCKS)
pdeb_pair_kernel(const __grid_constant__ PdebDevBatch b, const __grid_constant__ PdebConventions cv,
                 const __grid_constant__ PdebDevOut out, const int32_t* __restrict__ order, int64_t first, int64_t count) {
    static_assert(N % 2 == 0 && N >= 4 && N <= 62, "pair family: even meshes of 4..62 elements");
    constexpr int H = N / 2;
    typedef typename PdebVec2<real>::type real2;
    extern __shared__ __align__(16) unsigned char pdeb_smem[];
    const int tid = threadIdx.x;
    const int side = tid & 1;                                  /* 0: Dirichlet half, 1: free-end half */
    const int64_t t = ((int64_t)blockIdx.x * PDEB_PAIR_THREADS + tid) >> 1;
    const bool live = t < count;                               /* both lanes of a pair agree */
    const int64_t s = live ? (int64_t)order[first + t] : 0;
    const int64_t st = b.stride;
    real2* tab = reinterpret_cast<real2*>(pdeb_smem) + tid;    /* tab[(l-1) * T] = { R_ll, 1/pivot_l } */
    PdebClauseSlots<real> slots = pdeb_clause_slots<real>(pdeb_smem + sizeof(real2) * (size_t)PDEB_PAIR_THREADS * H,
                                                          PDEB_PAIR_THREADS, b.pmax, tid);
    constexpr int T = PDEB_PAIR_THREADS;

    int status = PDEB_OK;
    double dt = 1;
    int np = 0;
    int32_t S32 = 0;
    real w[H + 2];                        /* w[l], l = 1..H local rows; w[0] = 0 (beyond the outer end) */
    real vv[KIND == PDEB_KIND_WAVE ? H + 2 : 1], aa[KIND == PDEB_KIND_WAVE ? H + 2 : 1];
    real c[H];                            /* c[l], l = 1..H-1: multiplier of the edge (l, l+1) */
    real ro_o = 0, ro_i = 0;              /* R off-diagonal of an outer / inner material element */
    real k1 = 0, k2 = 0, k1o = 0;         /* middle 2x2 system: x_mine = k1 z_mine - k2 z_other, x_other = k1o z_other - k2 z_mine */
    double d_mid = 1.0, a_mid = 0.0;      /* pivot of the own middle row, coupling across the middle */
    real w_oth = 0, v_oth = 0, a_oth = 0; /* the partner's middle row */
    uint32_t emask = 0;                   /* bit m: local element m (between local rows m and m+1) is of the outer material */
    PdebPwl<real> in;
    in.init(nullptr, nullptr, 0, 0);
    const double* kt = b.knot_t + s;
    const double* ky = b.knot_y + s;
#pragma unroll
    for (int l = 0; l <= H + 1; ++l) w[l] = 0;
    if (KIND == PDEB_KIND_WAVE) {
#pragma unroll
        for (int l = 0; l <= H + 1; ++l) { vv[l] = 0; aa[l] = 0; }
    }

    if (live) {
        const double L = b.L[s], Tt = b.T[s], g0 = b.g0[s];
        dt = b.dt[s];
        const double x_a = b.x_a[s], x_b = b.x_b[s];
        const double rho_in = b.rho_in[s], rho_out = b.rho_out[s], E_in = b.E_in[s], E_out = b.E_out[s];
        np = b.n_pred[s];
        const int nk = b.n_knots[s];
        const int nt = b.n_tree[s];
        int64_t S = 0;
        if (!pdeb_params_ok(KIND, N, L, dt, Tt, g0, rho_in, rho_out, E_in, E_out)) status = PDEB_ST_BADPARAM;
        if (np < 1 || np > b.pmax || nk < 0 || nk > b.kmax || nt < 1 || nt > b.tmax) status = PDEB_ST_BADPARAM;
        if (status == PDEB_OK) {
            for (int j = 1; j < nk; ++j)
                if (!(kt[(int64_t)j * st] > kt[(int64_t)(j - 1) * st])) status = PDEB_ST_BADPARAM;
            for (int j = 0; j < np; ++j)
                if (b.pred_rel[(int64_t)j * st + s] > 2 || b.pred_top[(int64_t)j * st + s] > 1) status = PDEB_ST_BADPARAM;
        }
        if (status == PDEB_OK) {
            S = pdeb_step_count(Tt, dt, cv.steps_rule);
            if (S > PDEB_MAX_STEPS) status = PDEB_ST_BADPARAM;
        }
        const double h = L / N;
        if (status == PDEB_OK) {
            const int lumped = (KIND == PDEB_KIND_HEAT) ? cv.mass_heat : cv.mass_wave;
            const double cA = (KIND == PDEB_KIND_HEAT) ? cv.theta * dt : cv.newmark_beta * dt * dt;
            const double cR = (1.0 - cv.theta) * dt;
            /* local element m <-> global element: side 0: m, side 1: N - m (m = 0 does not exist there) */
#pragma unroll
            for (int m = 0; m <= H; ++m) {
                const int e = side ? (N - m) : m;
                const double xs = (cv.material_sampling == 1) ? e * h : (cv.material_sampling == 2) ? (e + 1) * h : (e + 0.5) * h;
                if (xs < x_a || xs > x_b) emask |= (1u << m);
            }
            const double m_o = rho_out * h, k_o = E_out / h, m_i = rho_in * h, k_i = E_in / h;
            const double mo_o = lumped ? 0.0 : m_o / 6.0, mo_i = lumped ? 0.0 : m_i / 6.0;
            const double md_o = lumped ? 0.5 * m_o : m_o / 3.0, md_i = lumped ? 0.5 * m_i : m_i / 3.0;
            const double ad_o = md_o + cA * k_o, ad_i = md_i + cA * k_i;
            const double ao_o = mo_o - cA * k_o, ao_i = mo_i - cA * k_i;
            const double rd_o = (KIND == PDEB_KIND_HEAT) ? md_o - cR * k_o : -k_o;
            const double rd_i = (KIND == PDEB_KIND_HEAT) ? md_i - cR * k_i : -k_i;
            ro_o = (real)((KIND == PDEB_KIND_HEAT) ? mo_o + cR * k_o : k_o);
            ro_i = (real)((KIND == PDEB_KIND_HEAT) ? mo_i + cR * k_i : k_i);
#define PDEB_EL(m, vo, vi) (((emask >> (m)) & 1u) ? (vo) : (vi))
            /* L D L^T from the outer end; local element 0 is missing on the free-end side */
            bool ok = true;
            double dprev = 0.0;
#pragma unroll
            for (int l = 1; l <= H; ++l) {
                const double has0 = (l == 1 && side) ? 0.0 : 1.0;
                double piv = has0 * PDEB_EL(l - 1, ad_o, ad_i) + PDEB_EL(l, ad_o, ad_i);
                if (l > 1) { const double aol = PDEB_EL(l - 1, ao_o, ao_i), ci = aol * dprev; c[l - 1] = (real)ci; piv -= ci * aol; }
                if (!(piv > 0.0)) ok = false;
                if (l == H) d_mid = piv;
                dprev = 1.0 / piv;
                real2 tv;
                tv.x = (real)(has0 * PDEB_EL(l - 1, rd_o, rd_i) + PDEB_EL(l, rd_o, rd_i));
                tv.y = (real)dprev;
                tab[(l - 1) * T] = tv;
            }
            a_mid = PDEB_EL(H, ao_o, ao_i);
            if (!isfinite(dprev)) ok = false;
            if (!ok) status = PDEB_ST_NONFINITE;
            in.init(kt, ky, nk, st);
            if (KIND == PDEB_KIND_WAVE) {
                /* d = 0 on the free nodes (deviation -g0), v = 0, a0 = M^-1 (f0 e_N - K w0) */
#pragma unroll
                for (int l = 1; l <= H; ++l) w[l] = (real)(-g0);
                w_oth = (real)(-g0);
            }
#undef PDEB_EL
        }
        if (status != PDEB_OK) { S = 0; np = 0; d_mid = 1.0; a_mid = 0.0; }
        int32_t k_end = 0;
        for (int j = 0; j < np; ++j) {
            const double m = b.pred_m[(int64_t)j * st + s], bb = b.pred_b[(int64_t)j * st + s];
            const double xlo = b.pred_xlo[(int64_t)j * st + s], xhi = b.pred_xhi[(int64_t)j * st + s];
            const double tlo = b.pred_tlo[(int64_t)j * st + s], thi = b.pred_thi[(int64_t)j * st + s];
            const double qlo = tlo / dt, qhi = thi / dt;
            double a0, a1;
            if (cv.window_rule == 1) { a0 = trunc(qlo); a1 = trunc(qhi); }
            else { a0 = ceil(qlo - PDEB_EPS); a1 = floor(qhi + PDEB_EPS); }
            if (!(a0 >= 0.0)) a0 = 0.0;
            if (!(a1 <= (double)S)) a1 = (double)S;
            if (a0 > a1) {
                double k = floor(0.5 * (tlo + thi) / dt + 0.5);
                if (!(k >= 0.0)) k = 0.0;
                if (k > (double)S) k = (double)S;
                a0 = a1 = k;
            }
            double ilo = ceil(xlo / h - PDEB_EPS), ihi = floor(xhi / h + PDEB_EPS);
            if (!(ilo >= 0.0)) ilo = 0.0;
            if (!(ihi <= (double)N)) ihi = (double)N;
            /* local node mask: bit l <-> node g(l) = side ? N + 1 - l : l, l = 0..H (l = 0 only on side 0) */
            uint32_t lm = 0;
            const bool none = !(ilo <= ihi);
            if (!none) {
                const int i0 = (int)ilo, i1 = (int)ihi;
#pragma unroll
                for (int l = 0; l <= H; ++l) {
                    const int g = side ? (N + 1 - l) : l;
                    if ((l > 0 || side == 0) && g >= i0 && g <= i1) lm |= (1u << l);
                }
            }
            const uint32_t top = b.pred_top[(int64_t)j * st + s], rel = b.pred_rel[(int64_t)j * st + s];
            const uint32_t endp = (none || cv.region_rule == 1) ? 1u : 0u;
            /* d = w_l - mh g - c0 with g = l (side 0) or N + 1 - l (side 1):  fma(smh, l, w_l) - c0s */
            const double mh = m * h;
            slots.R(j, 0) = (real)(side ? mh : -mh);
            slots.R(j, 1) = (real)((bb - g0) + (side ? mh * (double)(N + 1) : 0.0));
            slots.R(j, 2) = (top == PDEB_TOP_G) ? (real)INFINITY : (real)(-INFINITY);
            slots.W(j, 0) = (uint32_t)(int32_t)a0;
            slots.W(j, 1) = (uint32_t)(int32_t)a1;
            slots.W(j, 2) = lm;
            slots.W(j, 3) = rel | (top << 2) | (endp << 3) | (none ? 16u : 0u);
            if ((int32_t)a1 > k_end) k_end = (int32_t)a1;
        }
        S32 = k_end;
    }
    /* both lanes of a pair went through the same branches above (same sample); padding pairs idle */
    {
        /* middle 2x2: [d_mine a; a d_oth] (x_mine, x_oth) = (z_mine, z_oth); every lane of the warp takes part */
        const double d_oth = pdeb_partner(d_mid);
        const double det = d_mid * d_oth - a_mid * a_mid;
        k1 = (real)(d_oth / det);
        k1o = (real)(d_mid / det);
        k2 = (real)(a_mid / det);
        if (live && status == PDEB_OK && (!(det > 0.0) || !isfinite(1.0 / det))) { status = PDEB_ST_NONFINITE; S32 = 0; np = 0; }
    }

    auto sel_ro = [&](int m) -> real { return ((emask >> m) & 1u) ? ro_o : ro_i; };
    /* a' (wave) or w' (heat): solves A x = R xin + fN e_outer; xin[0] = 0, xin_oth = partner's middle row */
    auto solve = [&](const real* xin, real xin_oth, real fN, real* z, real& x_oth) {
        real zt = 0;
#pragma unroll
        for (int l = 1; l <= H; ++l) {
            const real2 tv = tab[(l - 1) * T];
            real r = tv.x * xin[l];
            if (l > 1) r = fma(sel_ro(l - 1), xin[l - 1], r);
            r = fma(sel_ro(l), (l < H) ? xin[l + 1] : xin_oth, r);
            if (l == 1) r += fN;
            if (l > 1) r = fma(-c[l - 1], zt, r);
            zt = r;
            if (l < H) z[l] = r * tv.y;
        }
        const real z_oth = pdeb_partner(zt);
        z[H] = fma(k1, zt, -k2 * z_oth);
        x_oth = fma(k1o, z_oth, -k2 * zt);
#pragma unroll
        for (int l = H - 1; l >= 1; --l) z[l] = fma(-c[l], z[l + 1], z[l]);
    };

    if (KIND == PDEB_KIND_WAVE) {
        if (__any_sync(PDEB_ALL, live && status == PDEB_OK)) {
            /* a0 = M^-1 (f0 e_N - K w0): R w0 through the same row formula, then the mass solve */
            const real f0 = in.eval(0.0, kt, ky, st);
            const int lumped = cv.mass_wave;
            const double L = live ? b.L[s] : 1.0;
            const double h = L / N, m_o = (live ? b.rho_out[s] : 1.0) * h, m_i = (live ? b.rho_in[s] : 1.0) * h;
            double r[H + 1];
#pragma unroll
            for (int l = 1; l <= H; ++l) {
                real rr = tab[(l - 1) * T].x * w[l];
                if (l > 1) rr = fma(sel_ro(l - 1), w[l - 1], rr);
                rr = fma(sel_ro(l), (l < H) ? w[l + 1] : w_oth, rr);
                if (l == 1 && side) rr += f0;
                r[l] = (double)rr;
            }
            if (lumped) {
#pragma unroll
                for (int l = 1; l <= H; ++l) {
                    const double ml = (l == 1 && side) ? 0.0 : (((emask >> (l - 1)) & 1u) ? m_o : m_i);
                    const double mr = ((emask >> l) & 1u) ? m_o : m_i;
                    aa[l] = (real)(r[l] / (0.5 * (ml + mr)));
                }
            } else {
                /* consistent mass: two-sided Thomas elimination with M, meeting in the middle like the step solve */
                double cp[H + 1], dprev = 0.0, zt = 0.0, dm = 1.0;
#pragma unroll
                for (int l = 1; l <= H; ++l) {
                    const double ml = (l == 1 && side) ? 0.0 : (((emask >> (l - 1)) & 1u) ? m_o : m_i);
                    const double mr = ((emask >> l) & 1u) ? m_o : m_i;
                    double piv = (ml + mr) / 3.0, rr = r[l];
                    if (l > 1) { const double mol = ml / 6.0, ci = mol * dprev; cp[l - 1] = ci; piv -= ci * mol; rr -= ci * zt; }
                    zt = rr;
                    if (l == H) dm = piv;
                    dprev = 1.0 / piv;
                    r[l] = rr * dprev;
                }
                const double amid = (((emask >> H) & 1u) ? m_o : m_i) / 6.0;
                const double d_oth = pdeb_partner(dm), z_oth = pdeb_partner(zt);
                const double det = dm * d_oth - amid * amid;
                r[H] = (zt * d_oth - amid * z_oth) / det;
#pragma unroll
                for (int l = H - 1; l >= 1; --l) r[l] -= cp[l] * r[l + 1];
#pragma unroll
                for (int l = 1; l <= H; ++l) aa[l] = (real)r[l];
            }
            a_oth = pdeb_partner(aa[H]);
            if (!live || status != PDEB_OK) {
#pragma unroll
                for (int l = 0; l <= H + 1; ++l) aa[l] = 0;
                a_oth = 0;
            }
        }
    }

    real f_prev = in.eval(0.0, kt, ky, st);
    int32_t next_ev = 0;
    uint32_t open = 0;

    for (int32_t k = 0;; ++k) {
        const bool on = (k <= S32);
        if (on && k == next_ev) {
            open = 0;
            next_ev = 0x7fffffff;
            for (int j = 0; j < np; ++j) {
                const int32_t a0 = (int32_t)slots.W(j, 0), a1 = (int32_t)slots.W(j, 1);
                if (k >= a0 && k <= a1) { open |= (1u << j); if (a1 + 1 < next_ev) next_ev = a1 + 1; }
                else if (k < a0 && a0 < next_ev) next_ev = a0;
            }
        }
        /* ---- (1) reduce the open clauses on the state of step k (both lanes of a pair take the same path) ---- */
        uint32_t todo = on ? open : 0u;
        while (__any_sync(PDEB_ALL, todo != 0u)) {
            const bool has = todo != 0u;
            int j = 0;
            if (has) { j = __ffs((int)todo) - 1; todo &= todo - 1u; }
            real smh = 0, c0s = 0, acc = 0;
            uint32_t lm = 0, info = 0;
            if (has) {
                smh = slots.R(j, 0);
                c0s = slots.R(j, 1);
                acc = slots.R(j, 2);
                lm = slots.W(j, 2);
                info = slots.W(j, 3);
            }
            const uint32_t rel = info & 3u;
            const real eq_eps = (cv.eq_rule == 1) ? (real)cv.eq_eps : (real)0;
            const uint32_t and_mask = (rel == PDEB_REL_EQ) ? 0x7fffffffu : 0xffffffffu;
            const uint32_t xor_mask = (rel == PDEB_REL_GT) ? 0x80000000u : 0u;
            real hi4[4];
            for (int q = 0; q < 4; ++q) hi4[q] = (real)(-INFINITY);
#pragma unroll
            for (int l = 0; l <= H; ++l) {
                const real d = ((l == 0) ? (real)0 : fma(smh, (real)l, w[l])) - c0s;
                const real g = pdeb_signop(d, and_mask, xor_mask);
                const bool tk = ((lm & (1u << l)) != 0u) & (g > hi4[l & 3]);
                hi4[l & 3] = tk ? g : hi4[l & 3];
            }
            const real a0 = hi4[0] > hi4[1] ? hi4[0] : hi4[1], a1 = hi4[2] > hi4[3] ? hi4[2] : hi4[3];
            real gmax = a0 > a1 ? a0 : a1;
            { const real o = pdeb_partner(gmax); gmax = o > gmax ? o : gmax; }
            real r = (rel == PDEB_REL_EQ) ? eq_eps - gmax : -gmax;
            if (__any_sync(PDEB_ALL, has && (info & 8u))) {
                /* uncommon: interpolated interval end points; each lane contributes the nodes it owns */
                double part[2] = { 0.0, 0.0 }, pr2[2] = { 0.0, 0.0 };
                const bool need = has && (info & 8u);
                if (need) {
                    const double L = b.L[s], h = L / N, rh = 1.0 / h;
                    const double xe[2] = { b.pred_xlo[(int64_t)j * st + s], b.pred_xhi[(int64_t)j * st + s] };
                    const double m = b.pred_m[(int64_t)j * st + s], bb = b.pred_b[(int64_t)j * st + s];
                    real wl[H + 2];
#pragma unroll
                    for (int l = 0; l <= H; ++l) wl[l] = (l == 0) ? (real)0 : w[l];
                    for (int q = 0; q < 2; ++q) {
                        const double xx = xe[q] < 0 ? 0 : (xe[q] > L ? L : xe[q]);
                        int e = (int)(xx * rh);
                        if (e < 0) e = 0;
                        if (e > N - 1) e = N - 1;
                        const double xi = (xx - e * h) * rh;
                        /* nodes e (weight 1 - xi) and e + 1 (weight xi); local index of node g: side ? N + 1 - g : g */
                        double acc2 = 0.0;
                        for (int d2 = 0; d2 < 2; ++d2) {
                            const int g = e + d2;
                            const int l = side ? (N + 1 - g) : g;
                            const bool mine = side ? (l >= 1 && l <= H) : (l >= 0 && l <= H);
                            if (mine) acc2 += (d2 ? xi : 1.0 - xi) * (double)wl[l];
                        }
                        part[q] = acc2;
                        pr2[q] = m * xx + bb;
                    }
                }
                const double g0 = live ? b.g0[s] : 0.0;
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const double u = g0 + (part[q] + pdeb_partner(part[q]));
                    if (need) {
                        double mg;
                        if (rel == PDEB_REL_GT) mg = u - pr2[q];
                        else if (rel == PDEB_REL_LT) mg = pr2[q] - u;
                        else { const double lo = u - (pr2[q] - (double)eq_eps), hi = (pr2[q] + (double)eq_eps) - u; mg = lo < hi ? lo : hi; }
                        if (q == 0 && (info & 16u)) r = (real)mg;
                        else if ((real)mg < r) r = (real)mg;
                    }
                }
            }
            if (has) {
                const bool take = (((info >> 2) & 1u) == PDEB_TOP_G) ? (r < acc) : (r > acc);
                if (take || r != r) slots.R(j, 2) = r;
            }
        }
        if (!__any_sync(PDEB_ALL, k < S32)) break;
        /* ---- (2) advance to step k+1 ---- */
        {
            const real f_next = in.eval((double)(k + 1) * dt, kt, ky, st);
            if (KIND == PDEB_KIND_HEAT) {
                const real th_dt = (real)(cv.theta * dt);
                const real fb = side ? fma((real)dt - th_dt, f_prev, th_dt * f_next) : (real)0;
                real z[H + 2];
                real x_oth;
                solve(w, w_oth, fb, z, x_oth);
#pragma unroll
                for (int l = 1; l <= H; ++l) w[l] = z[l];
                w_oth = x_oth;
            } else {
                const real cw1 = (real)dt;
                const real cw2 = (real)(0.5 * (1.0 - 2.0 * cv.newmark_beta) * dt * dt);
                const real cw3 = (real)((1.0 - cv.newmark_gamma) * dt);
#pragma unroll
                for (int l = 1; l <= H; ++l) {
                    w[l] = fma(cw2, aa[l], fma(cw1, vv[l], w[l]));
                    vv[l] = fma(cw3, aa[l], vv[l]);
                }
                w_oth = fma(cw2, a_oth, fma(cw1, v_oth, w_oth));
                v_oth = fma(cw3, a_oth, v_oth);
                real x_oth;
                solve(w, w_oth, side ? f_next : (real)0, aa, x_oth);
                const real cw4 = (real)(cv.newmark_beta * dt * dt);
                const real cw5 = (real)(cv.newmark_gamma * dt);
#pragma unroll
                for (int l = 1; l <= H; ++l) {
                    w[l] = fma(cw4, aa[l], w[l]);
                    vv[l] = fma(cw5, aa[l], vv[l]);
                }
                a_oth = x_oth;
                w_oth = fma(cw4, x_oth, w_oth);
                v_oth = fma(cw5, x_oth, v_oth);
            }
            f_prev = f_next;
        }
    }

    if (!live || side) return;
    /* ---- formula tree and outputs (lane A of the pair) ---- */
    double val[PDEB_MAX_TREE];
    double rob = PDEB_FAILED_ROBUSTNESS;
    if (status == PDEB_OK) {
        const int nt = b.n_tree[s];
        for (int j = 0; j < nt; ++j) {
            const int op = b.tree_op[(int64_t)j * st + s], lh = b.tree_lhs[(int64_t)j * st + s], rh = b.tree_rhs[(int64_t)j * st + s];
            double x = NAN;
            if (op == PDEB_TREE_LEAF) { if (lh >= 0 && lh < np) x = (double)slots.R(lh, 2); else status = PDEB_ST_BADPARAM; }
            else if ((op == PDEB_TREE_AND || op == PDEB_TREE_OR) && lh >= 0 && lh < j && rh >= 0 && rh < j) {
                const double a0 = val[lh], a1 = val[rh];
                if (a0 != a0 || a1 != a1) x = NAN;
                else if (op == PDEB_TREE_AND) x = a0 < a1 ? a0 : a1;
                else x = a0 > a1 ? a0 : a1;
            } else status = PDEB_ST_BADPARAM;
            val[j] = x;
        }
        if (status == PDEB_OK) {
            rob = val[nt - 1];
            if (!isfinite(rob)) status = PDEB_ST_NONFINITE;
            for (int j = 0; j < np; ++j) if (!isfinite((double)slots.R(j, 2))) status = PDEB_ST_NONFINITE;
        }
    }
    out.robustness[s] = (status == PDEB_OK) ? rob : PDEB_FAILED_ROBUSTNESS;
    if (out.status) out.status[s] = status;
    if (out.clause) {
        for (int j = 0; j < b.pmax; ++j)
            out.clause[(int64_t)j * st + s] = (status == PDEB_OK && j < np) ? (double)slots.R(j, 2) : PDEB_FAILED_ROBUSTNESS;
    }
}

template <int KIND, int N, typename real>
inline size_t pdeb_pair_smem_bytes(int pmax) {
    return sizeof(typename PdebVec2<real>::type) * (size_t)PDEB_PAIR_THREADS * (N / 2) +
           pdeb_clause_smem_bytes<real>(PDEB_PAIR_THREADS, pmax);
}

#endif /* PDEB_PAIR_CUH */

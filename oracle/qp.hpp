// TEST INFRASTRUCTURE -- part of the CPU oracle (see oracle/README.md).  Never linked into the product.
//
// Dense QP used by the oracle in place of qpOASES (third-party, `saga0619/qpoases` fork, version
// unpinned by the reference -- install_prereq.sh:37; source not available offline).
//
// What is restated: the *result* the reference obtains from
//     CQuadraticProgram::SolveQPoases            (reference src/qp_wrapper.cpp:192-380)
// for the QPs it builds in CalcSingleTaskTorqueWithQP (reference src/dwbc.cpp:982-1053) and
// CalcContactRedistribute (reference src/dwbc.cpp:1458-1546):
//     min 1/2 x^T H x   s.t.  A x <= ubA          (g = 0, lbA = -1e20, no bounds on x)
// with H = diag(I_t, 0_r) (task QP) or H = I (redistribution QP).  H is only positive
// SEMI-definite for the task QP.  qpOASES' setToMPC() options regularise the Hessian with
// eps = 1e3*EPS and run one proximal correction step, so the solution it returns is the
// LEXICOGRAPHIC optimum:  (1) minimise ||x_t||^2, (2) among those minimisers take the one
// with minimum ||x_c||^2.  SURVEY.md section 7 (hard part 1) verified this on the reference's
// golden vectors; tests/test_oracle_golden.py re-verifies it (torque_task_ / torque_contact_
// of tests/cases/{1,2}).
//
// Algorithm here (chosen to be independent from the product's Goldfarb-Idnani GPU solver):
//   * strictly convex sub-problems are solved as least-distance problems through
//     Lawson-Hanson NNLS (Lawson & Hanson, "Solving Least Squares Problems", ch. 23),
//   * stage 1 (min ||x_t||^2 with x_c free) uses the proximal-point iteration
//         x^{k+1} = argmin 1/2||x_t||^2 + delta/2 ||x_c - x_c^k||^2  s.t. A x <= ub,
//     which converges linearly with a ratio ~delta/(delta+sigma) << 1,
//   * stage 2 minimises ||x_c||^2 on the optimal face of stage 1: the strongly active rows
//     become equalities, eliminated with a rank-revealing QR, and the remaining rows form
//     another least-distance problem in the leftover directions.
#pragma once
#include <cstdio>
#include <cstdlib>
#include "linalg.hpp"

namespace orc {

static const double LDP_VIOL_TOL = 1e-9;

// Householder least squares: minimise ||E z - f||, E (m x p) full column rank, m >= p.
inline bool ls_solve(const Mat &E, const std::vector<double> &f, std::vector<double> &z) {
    int m = E.r, p = E.c;
    Mat R = E;
    std::vector<double> b = f;
    for (int k = 0; k < p; k++) {
        double tau, beta;
        make_householder(&R(k, k), m - k, tau, beta);
        for (int j = k + 1; j < p; j++) {
            double s = R(k, j);
            for (int i = k + 1; i < m; i++) s += R(i, k) * R(i, j);
            s *= tau;
            R(k, j) -= s;
            for (int i = k + 1; i < m; i++) R(i, j) -= s * R(i, k);
        }
        double s = b[k];
        for (int i = k + 1; i < m; i++) s += R(i, k) * b[i];
        s *= tau;
        b[k] -= s;
        for (int i = k + 1; i < m; i++) b[i] -= s * R(i, k);
        R(k, k) = beta;
    }
    z.assign(p, 0.0);
    for (int i = p - 1; i >= 0; i--) {
        double v = b[i];
        for (int k = i + 1; k < p; k++) v -= R(i, k) * z[k];
        if (R(i, i) == 0.0) return false;
        z[i] = v / R(i, i);
    }
    return true;
}

// Lawson-Hanson NNLS: min ||E x - f||, x >= 0.
// viol_tol (LDP mode): E = [G^T; h^T] with unit-norm G rows and f = e_last; then w_j / res_last is the
// violation of constraint j at the current candidate point, and constraint j only enters while that
// violation exceeds viol_tol[j].
inline void nnls(const Mat &E, const std::vector<double> &f, std::vector<double> &x, int max_iter = 0,
                 const std::vector<double> *viol_tol = nullptr) {
    int m = E.r, n = E.c;
    if (max_iter <= 0) max_iter = 10 * n + 50;
    x.assign(n, 0.0);
    std::vector<char> inP(n, 0);
    std::vector<double> w(n), res(m), z;
    std::vector<double> cn(n);
    for (int j = 0; j < n; j++) {
        double s = 0;
        for (int i = 0; i < m; i++) s += E(i, j) * E(i, j);
        cn[j] = std::sqrt(s);
    }
    double fn = 0;
    for (double v : f) fn += v * v;
    fn = std::sqrt(fn);
    std::vector<char> banned(n, 0);
    for (int it = 0; it < max_iter; it++) {
        for (int i = 0; i < m; i++) {
            double s = f[i];
            for (int j = 0; j < n; j++)
                if (x[j] != 0.0) s -= E(i, j) * x[j];
            res[i] = s;
        }
        int t = -1;
        double best = 0;
        for (int j = 0; j < n; j++) {
            if (inP[j] || banned[j] || cn[j] == 0) continue;
            double s = 0;
            for (int i = 0; i < m; i++) s += E(i, j) * res[i];
            w[j] = s;
            // dual feasibility test relative to the column and right-hand-side scales
            double thr = 1e-13 * cn[j] * fn;
            if (viol_tol) thr = std::max(thr, (*viol_tol)[j] * res[m - 1]);
            if (s > thr && s / cn[j] > best) best = s / cn[j], t = j;
        }
        if (t < 0) break;
        inP[t] = 1;
        bool first = true;
        while (true) {
            std::vector<int> P;
            for (int j = 0; j < n; j++)
                if (inP[j]) P.push_back(j);
            if (P.empty()) break;
            Mat EP(m, (int)P.size());
            for (size_t k = 0; k < P.size(); k++)
                for (int i = 0; i < m; i++) EP(i, (int)k) = E(i, P[k]);
            if ((int)P.size() > m || !ls_solve(EP, f, z)) {  // dependent column: reject it
                inP[t] = 0;
                banned[t] = 1;
                break;
            }
            if (first) {
                // Lawson-Hanson safeguard: the entering variable must come out positive
                size_t kt = 0;
                for (size_t k = 0; k < P.size(); k++)
                    if (P[k] == t) kt = k;
                if (z[kt] <= 0) {
                    inP[t] = 0;
                    banned[t] = 1;
                    break;
                }
                first = false;
            }
            bool allpos = true;
            for (size_t k = 0; k < P.size(); k++)
                if (z[k] <= 0) allpos = false;
            if (allpos) {
                for (int j = 0; j < n; j++) x[j] = 0;
                for (size_t k = 0; k < P.size(); k++) x[P[k]] = z[k];
                std::fill(banned.begin(), banned.end(), 0);
                break;
            }
            double alpha = 2.0;
            for (size_t k = 0; k < P.size(); k++)
                if (z[k] <= 0) {
                    double a = x[P[k]] / (x[P[k]] - z[k]);
                    if (a < alpha) alpha = a;
                }
            for (size_t k = 0; k < P.size(); k++) {
                int j = P[k];
                x[j] += alpha * (z[k] - x[j]);
            }
            for (size_t k = 0; k < P.size(); k++) {
                int j = P[k];
                if (z[k] <= 0 && x[j] <= 1e-300 + 1e-15 * std::fabs(z[k])) {
                    x[j] = 0;
                    inP[j] = 0;
                }
            }
        }
    }
}

// Least-distance programming: min ||x|| s.t. G x >= h.  Returns false if infeasible.
// lambda (size m) receives the KKT multipliers, x = G^T lambda.
// row_tol (optional): violation tolerance of each row in the units of the rows as passed in; default
// LDP_VIOL_TOL * max(||G_i||, |h_i|).
inline bool ldp(const Mat &G, const std::vector<double> &h, std::vector<double> &x, std::vector<double> &lambda,
                const std::vector<double> *row_tol = nullptr) {
    int m = G.r, n = G.c;
    x.assign(n, 0.0);
    lambda.assign(m, 0.0);
    // row scaling (does not change the feasible set); all-zero rows are decided directly
    // A row whose norm is below 1e-10 of the largest row norm is numerically zero: in the hierarchy such rows are what is
    // left of a constraint that is already active at a higher level (its normal lies in the range the null-space projector
    // removed, so only rounding noise remains).  Normalising it would turn noise into a constraint; qpOASES discards such
    // rows through its epsNum/epsDen ratio tests.
    std::vector<int> rows;
    std::vector<double> scale, vtol;
    double smax = 0;
    for (int i = 0; i < m; i++) {
        double s = 0;
        for (int j = 0; j < n; j++) s += G(i, j) * G(i, j);
        smax = std::max(smax, std::sqrt(s));
    }
    double habs = 0;  // magnitude of the right-hand sides: the feasibility margin of a numerically zero row is relative to it
    for (int i = 0; i < m; i++) habs = std::max(habs, std::fabs(h[i]));
    for (int i = 0; i < m; i++) {
        double s = 0;
        for (int j = 0; j < n; j++) s += G(i, j) * G(i, j);
        s = std::sqrt(s);
        if (s < 1e-14 || s <= 1e-10 * smax) {
            if (h[i] > 1e-9 * (1.0 + habs)) return false;
            continue;
        }
        rows.push_back(i);
        scale.push_back(1.0 / s);
        vtol.push_back((row_tol ? (*row_tol)[i] : LDP_VIOL_TOL * std::max(s, std::fabs(h[i]))) / s);
    }
    int mr = (int)rows.size();
    if (mr == 0) return true;
    bool need = false;
    for (int k = 0; k < mr; k++)
        if (h[rows[k]] > 0) need = true;
    if (!need) return true;  // x = 0 is feasible, hence optimal
    // NNLS runs in "LDP mode": a constraint only enters while it is violated by more than
    // LDP_VIOL_TOL (relative to its own scale).  The hierarchy makes lower-level QPs inherit the
    // active rows of the levels above, so their feasible sets are routinely lower-dimensional
    // faces known only to rounding; without a violation tolerance the feasibility decision is a
    // coin flip.  qpOASES does the same through its bound tolerances.
    Mat E(n + 1, mr);
    for (int k = 0; k < mr; k++) {
        for (int j = 0; j < n; j++) E(j, k) = G(rows[k], j) * scale[k];
        E(n, k) = h[rows[k]] * scale[k];
    }
    std::vector<double> f(n + 1, 0.0), u;
    f[n] = 1.0;
    nnls(E, f, u, 0, &vtol);
    std::vector<double> r(n + 1, 0.0);
    for (int i = 0; i <= n; i++) {
        double s = -f[i];
        for (int k = 0; k < mr; k++) s += E(i, k) * u[k];
        r[i] = s;
    }
    double rn = 0;
    for (double v : r) rn += v * v;
    rn = std::sqrt(rn);
    if (getenv("ORC_QP_DEBUG")) fprintf(stderr, "  ldp: n=%d mr=%d rn=%.3e r[n]=%.3e\n", n, mr, rn, r[n]);
    if (rn < 1e-11 || !(r[n] < 0)) return false;
    for (int j = 0; j < n; j++) x[j] = -r[j] / r[n];
    // Guard against a false "feasible": for an infeasible problem the NNLS residual is zero up to rounding, and a residual that
    // lands just above the threshold gives a meaningless x = -r / r[n] (seen once in 65 536 states: |r| = 3e-11, rows violated by
    // 49).  The answer must satisfy every row to the accuracy of this solver.
    {
        double xn2 = 0;
        for (int j = 0; j < n; j++) xn2 += x[j] * x[j];
        const double xn = std::sqrt(xn2);
        for (int k = 0; k < mr; k++) {
            double gx = 0;
            for (int j = 0; j < n; j++) gx += G(rows[k], j) * x[j];
            const double viol = (h[rows[k]] - gx) * scale[k];  // > 0: violated, in units of the normalised row
            if (viol > 1e-6 * (1.0 + xn + std::fabs(h[rows[k]]) * scale[k])) return false;
        }
    }
    for (int k = 0; k < mr; k++) lambda[rows[k]] = u[k] * scale[k] / (-r[n]);
    return true;
}

struct QpResult {
    int status = 0;  // 1 solved, 0 infeasible / failed
    std::vector<double> x;
    std::vector<int> active;  // rows with A x = ub at the solution (|slack| <= tol)
    std::vector<int> gray;    // rows whose relative slack lies within a decade of that threshold, (-1e-8, -1e-10): weakly active, may be reported either way
    int prox_iters = 0;
};

// min 1/2 ||x||^2 s.t. A x <= ub
inline QpResult qp_min_norm(const Mat &A, const std::vector<double> &ub, const std::vector<double> *row_tol = nullptr) {
    QpResult res;
    Mat G = -A;
    std::vector<double> h(ub.size()), lam;
    for (size_t i = 0; i < ub.size(); i++) h[i] = -ub[i];
    if (!ldp(G, h, res.x, lam, row_tol)) return res;
    res.status = 1;
    // Polish: the NNLS-based LDP lets a row in only while it is violated by more than its tolerance, so the rows of the optimal
    // face are met to that tolerance (1e-8 of the row scale in the loose stages) rather than exactly.  A lower level of the hierarchy
    // then sees those rows with a right-hand side of +-1e-8 instead of 0 and "is this row active" becomes a coin flip.  With the
    // multipliers identifying the face W, the answer is re-computed as the least-norm point of A_W x = ub_W (the same point, to
    // rounding) and kept when every other row still holds.
    {
        double lmax = 0;
        for (double v : lam) lmax = std::max(lmax, v);
        std::vector<int> W;
        for (int i = 0; i < A.r; i++)
            if (lam[i] > 1e-9 * lmax && lmax > 0) W.push_back(i);
        const int n = A.c, q = (int)W.size();
        if (q > 0 && q <= n) {
            Mat AW(q, n);
            for (int k = 0; k < q; k++)
                for (int j = 0; j < n; j++) AW(k, j) = A(W[k], j);
            Cod f = cod_compute(AW, 1e-11);
            if (f.rank == q) {
                std::vector<double> xp(n, 0.0);
                for (int j = 0; j < n; j++)
                    for (int k = 0; k < q; k++) xp[j] += f.pinv(j, k) * ub[W[k]];
                double dn = 0, xn = 0;
                for (int j = 0; j < n; j++) dn += (xp[j] - res.x[j]) * (xp[j] - res.x[j]), xn += xp[j] * xp[j];
                bool ok = std::sqrt(dn) <= 1e-6 * (1.0 + std::sqrt(xn));
                for (int i = 0; i < A.r && ok; i++) {
                    double s = -ub[i], nn = 0;
                    for (int j = 0; j < n; j++) s += A(i, j) * xp[j], nn += A(i, j) * A(i, j);
                    const double tol = row_tol ? (*row_tol)[i] : LDP_VIOL_TOL * std::max(std::sqrt(nn), std::fabs(ub[i]));
                    if (s > tol) ok = false;
                }
                if (ok) res.x = xp;
            }
        }
    }
    return res;
}

inline void fill_active(const Mat &A, const std::vector<double> &ub, QpResult &res) {
    res.active.clear();
    res.gray.clear();
    for (int i = 0; i < A.r; i++) {
        double s = -ub[i], nn = 0;
        for (int j = 0; j < A.c; j++) s += A(i, j) * res.x[j], nn += A(i, j) * A(i, j);
        if (nn == 0) continue;
        if (s > -1e-9 * (1.0 + std::fabs(ub[i]))) res.active.push_back(i);
        const double rel = s / (1.0 + std::fabs(ub[i]));
        if (rel > -1e-8 && rel < -1e-10) res.gray.push_back(i);
    }
}

// Lexicographic QP: variables x = [x_t (t); x_c (r)], minimise ||x_t||^2 then ||x_c||^2
// subject to A x <= ub.
inline QpResult lexqp(const Mat &A, const std::vector<double> &ub, int t, double delta0 = 1e-5) {
    QpResult res;
    int m = A.r, n = A.c, r = n - t;
    res.x.assign(n, 0.0);
    if (r == 0) {
        res = qp_min_norm(A, ub);
        if (res.status) fill_active(A, ub, res);
        return res;
    }
    Mat At = A.left_cols(t), Ac = A.right_cols(r);
    // violation tolerances in the units of the original rows: tight for stage 1, looser for the
    // (inherently thin) stage-2 sets so that stage-1 rounding cannot make them "infeasible"
    std::vector<double> rtol(m), rtol2(m);
    for (int i = 0; i < m; i++) {
        double s = 0;
        for (int j = 0; j < n; j++) s += A(i, j) * A(i, j);
        double sc = std::max(std::sqrt(s), std::fabs(ub[i]));
        rtol[i] = 1e-10 * sc;
        rtol2[i] = 1e-8 * sc;
    }
    // Step A: x_t = 0 attainable?  Then the answer is the min-norm x_c.
    {
        QpResult s2 = qp_min_norm(Ac, ub, &rtol2);
        if (s2.status) {
            for (int j = 0; j < r; j++) res.x[t + j] = s2.x[j];
            res.status = 1;
            fill_active(A, ub, res);
            return res;
        }
    }
    if (t == 0) return res;
    // Stage 1: proximal-point iteration.
    double nt = 0, nc = 0;
    for (int j = 0; j < t; j++)
        for (int i = 0; i < m; i++) nt += At(i, j) * At(i, j);
    for (int j = 0; j < r; j++)
        for (int i = 0; i < m; i++) nc += Ac(i, j) * Ac(i, j);
    double delta = delta0 * (nc / r) / std::max(nt / t, 1e-300);
    double sq = std::sqrt(delta);
    Mat G(m, n);
    auto scale_G = [&]() {
        for (int i = 0; i < m; i++) {
            for (int j = 0; j < t; j++) G(i, j) = -At(i, j);
            for (int j = 0; j < r; j++) G(i, t + j) = -Ac(i, j) / sq;
        }
    };
    scale_G();
    std::vector<double> c(r, 0.0), h(m), y, lam, xprev(n, 0.0), x(n, 0.0);
    bool conv = false;
    double dn_prev = 0.0;
    int reductions = 0;
    for (int k = 0; k < 40; k++) {
        for (int i = 0; i < m; i++) {
            double s = ub[i];
            for (int j = 0; j < r; j++) s -= Ac(i, j) * c[j];
            h[i] = -s;
        }
        if (!ldp(G, h, y, lam, &rtol)) return res;
        for (int j = 0; j < t; j++) x[j] = y[j];
        for (int j = 0; j < r; j++) x[t + j] = c[j] + y[t + j] / sq;
        res.prox_iters = k + 1;
        // convergence is judged on x_t only: x_c is recomputed exactly in stage 2
        double dn = 0, xn = 0;
        for (int j = 0; j < t; j++) dn += (x[j] - xprev[j]) * (x[j] - xprev[j]), xn += x[j] * x[j];
        xprev = x;
        for (int j = 0; j < r; j++) c[j] = x[t + j];
        if (getenv("ORC_QP_DEBUG")) fprintf(stderr, "prox %d dx_t %.3e |x_t| %.3e delta %.3e\n", k, std::sqrt(dn), std::sqrt(xn), delta);
        // converged; or stalled at the accuracy of the NNLS-based inner solve (seen at |x_t| ~ 1e3: steps of 1e-10 relative that
        // never reach 1e-12) -- the iterate is then as good as this oracle gets and far inside the 1e-6 parity tolerance
        if (k > 0 && (std::sqrt(dn) <= 1e-12 * (1.0 + std::sqrt(xn)) || (k >= 8 && std::sqrt(dn) <= 5e-10 * (1.0 + std::sqrt(xn))))) {
            conv = true;
            break;
        }
        // contraction ratio delta / (delta + sigma): on a nearly degenerate optimal face (sigma tiny) the default delta crawls;
        // a step that shrinks by less than half switches to a 100 times smaller delta.  The limit does not depend on delta.
        if (k >= 2 && reductions < 3 && dn > 0.25 * dn_prev) {
            delta *= 0.01;
            sq = std::sqrt(delta);
            scale_G();
            reductions++;
        }
        dn_prev = dn;
    }
    if (!conv) return res;
    // Stage 2 on the optimal face: rows with a positive multiplier are equalities.
    double lmax = 0;
    for (double v : lam) lmax = std::max(lmax, v);
    std::vector<int> W, rest;
    for (int i = 0; i < m; i++) (lam[i] > 1e-9 * lmax ? W : rest).push_back(i);
    int q = (int)W.size();
    std::vector<double> xt(x.begin(), x.begin() + t);
    Mat Et(r, q);  // E^T, E = A_c[W]
    std::vector<double> e(q);
    for (int k = 0; k < q; k++) {
        double s = ub[W[k]];
        for (int j = 0; j < t; j++) s -= At(W[k], j) * xt[j];
        e[k] = s;
        for (int j = 0; j < r; j++) Et(j, k) = Ac(W[k], j);
    }
    Cod f = cod_compute(Et, 1e-9);  // E^T P = Q R
    int rho = f.rank;
    if (getenv("ORC_QP_DEBUG")) fprintf(stderr, "stage2: q=%d rho=%d rest=%d lmax=%g\n", q, rho, (int)rest.size(), lmax);
    // R[0:rho,0:rho]^T y1 = (P^T e)[0:rho]
    std::vector<double> y1(rho, 0.0);
    for (int i = 0; i < rho; i++) {
        double v = e[f.perm[i]];
        for (int k = 0; k < i; k++) v -= f.R(k, i) * y1[k];
        y1[i] = v / f.R(i, i);
    }
    std::vector<double> xc(r, 0.0);
    for (int j = 0; j < r; j++)
        for (int i = 0; i < rho; i++) xc[j] += f.Q(j, i) * y1[i];
    if (rho < r && !rest.empty()) {
        int fr = r - rho, mr = (int)rest.size();
        Mat C2(mr, fr);
        std::vector<double> d2(mr), tol2(mr);
        for (int k = 0; k < mr; k++) {
            int i = rest[k];
            tol2[k] = rtol2[i];
            double s = ub[i];
            for (int j = 0; j < t; j++) s -= At(i, j) * xt[j];
            for (int j = 0; j < r; j++) s -= Ac(i, j) * xc[j];
            d2[k] = s;
            for (int l = 0; l < fr; l++) {
                double v = 0;
                for (int j = 0; j < r; j++) v += Ac(i, j) * f.Q(j, rho + l);
                C2(k, l) = v;
            }
        }
        QpResult s2 = qp_min_norm(C2, d2, &tol2);
        if (getenv("ORC_QP_DEBUG")) fprintf(stderr, "stage2 ldp status %d\n", s2.status);
        if (!s2.status) return res;
        for (int j = 0; j < r; j++)
            for (int l = 0; l < fr; l++) xc[j] += f.Q(j, rho + l) * s2.x[l];
    }
    for (int j = 0; j < t; j++) res.x[j] = xt[j];
    for (int j = 0; j < r; j++) res.x[t + j] = xc[j];
    res.status = 1;
    fill_active(A, ub, res);
    return res;
}

}  // namespace orc

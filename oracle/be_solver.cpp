// ORACLE — TEST INFRASTRUCTURE ONLY.  See be_solver.h.
#include "be_solver.h"

#include <cfloat>
#include <cmath>

namespace orc {

void Problem::AddParameterBlock(double* ptr, int size, bool is_pose) {
    if (find(ptr) >= 0) return;
    Block b;
    b.ptr = ptr;
    b.size = size;
    b.local = is_pose ? 6 : size;
    b.is_pose = is_pose;
    blocks.push_back(b);
}
int Problem::find(double* ptr) const {
    for (size_t i = 0; i < blocks.size(); i++)
        if (blocks[i].ptr == ptr) return (int)i;
    return -1;
}
void Problem::SetParameterBlockConstant(double* ptr) { blocks[find(ptr)].constant = true; }
void Problem::AddResidualBlock(std::shared_ptr<CostFunction> cf, const CauchyLoss* loss, std::vector<double*> params) {
    Res r;
    r.cf = cf;
    r.loss = loss;
    for (size_t i = 0; i < params.size(); i++) {
        int k = find(params[i]);
        if (k < 0) {
            AddParameterBlock(params[i], cf->block_sizes[i], false);
            k = (int)blocks.size() - 1;
        }
        blocks[k].used = true;
        r.blocks.push_back(k);
    }
    residuals.push_back(r);
}

namespace {

struct ResEval {
    std::vector<double> r;          // corrected residuals
    std::vector<Mat> J;             // corrected, local-size columns, scaled (empty for constant blocks)
};

struct Program {
    Problem* prob;
    std::vector<int> active;        // indices of non-constant, used blocks in column order (others first, landmarks last)
    int n_other = 0, n_land = 0;    // local dimensions
    int ncols = 0;
    std::vector<bool> is_landmark;  // per block
};

// Landmarks (the e-blocks of DENSE_SCHUR): size-1 blocks that never share a residual block with
// another already chosen size-1 block (Ceres picks an independent set; with VINS's factor graph that is
// the inverse depths, td being adjacent to all of them).
Program reduce_program(Problem& p) {
    Program g;
    g.prob = &p;
    const int nb = (int)p.blocks.size();
    std::vector<int> degree(nb, 0);
    for (auto& r : p.residuals)
        for (int b : r.blocks) degree[b]++;
    g.is_landmark.assign(nb, false);
    std::vector<int> cand;
    for (int b = 0; b < nb; b++)
        if (p.blocks[b].size == 1 && p.blocks[b].used && !p.blocks[b].constant) cand.push_back(b);
    std::stable_sort(cand.begin(), cand.end(), [&](int a, int b) { return degree[a] < degree[b]; });
    for (int b : cand) {
        bool independent = true;
        for (auto& r : p.residuals) {
            bool has_b = false, has_other = false;
            for (int k : r.blocks) {
                if (k == b) has_b = true;
                else if (g.is_landmark[k]) has_other = true;
            }
            if (has_b && has_other) {
                independent = false;
                break;
            }
        }
        if (independent) g.is_landmark[b] = true;
    }
    int off = 0;
    for (int b = 0; b < nb; b++)
        if (p.blocks[b].used && !p.blocks[b].constant && !g.is_landmark[b]) {
            p.blocks[b].offset = off;
            off += p.blocks[b].local;
            g.active.push_back(b);
        }
    g.n_other = off;
    for (int b = 0; b < nb; b++)
        if (g.is_landmark[b]) {
            p.blocks[b].offset = off;
            off += 1;
            g.active.push_back(b);
        }
    g.n_land = off - g.n_other;
    g.ncols = off;
    return g;
}

// Residuals (+ Jacobians) of every block with the robust corrector applied, as Ceres' ResidualBlock::Evaluate.
double evaluate(const Program& g, const std::vector<std::vector<double>>& x, bool want_jac, std::vector<ResEval>* out) {
    Problem& p = *g.prob;
    double cost = 0;
    if (out) out->assign(p.residuals.size(), ResEval());
    std::vector<const double*> params;
    for (size_t ri = 0; ri < p.residuals.size(); ri++) {
        const Problem::Res& rb = p.residuals[ri];
        const int nr = rb.cf->num_residuals, nbk = (int)rb.blocks.size();
        params.resize(nbk);
        for (int i = 0; i < nbk; i++) params[i] = x[rb.blocks[i]].data();
        std::vector<double> r(nr);
        std::vector<Mat> Jg;
        std::vector<double*> raw(nbk, nullptr);
        if (want_jac) {
            for (int i = 0; i < nbk; i++) Jg.emplace_back(nr, rb.cf->block_sizes[i]);
            for (int i = 0; i < nbk; i++) {
                const Problem::Block& b = p.blocks[rb.blocks[i]];
                raw[i] = (b.constant || !b.used) ? nullptr : Jg[i].d.data();
            }
        }
        rb.cf->Evaluate(params.data(), r.data(), want_jac ? raw.data() : nullptr);
        double sq_norm = 0;
        for (double v : r) sq_norm += v * v;
        if (!rb.loss) {
            cost += 0.5 * sq_norm;
        } else {
            double rho[3];
            rb.loss->Evaluate(sq_norm, rho);
            cost += 0.5 * rho[0];
            if (out) {  // Corrector
                const double sqrt_rho1 = std::sqrt(rho[1]);
                double residual_scaling, alpha_sq_norm;
                if ((sq_norm == 0.0) || (rho[2] <= 0.0)) {
                    residual_scaling = sqrt_rho1;
                    alpha_sq_norm = 0.0;
                } else {
                    const double D = 1.0 + 2.0 * sq_norm * rho[2] / rho[1];
                    const double alpha = 1.0 - std::sqrt(D);
                    residual_scaling = sqrt_rho1 / (1 - alpha);
                    alpha_sq_norm = alpha / sq_norm;
                }
                if (want_jac)
                    for (int i = 0; i < nbk; i++) {
                        if (!raw[i]) continue;
                        Mat& J = Jg[i];
                        std::vector<double> rtJ(J.c, 0.0);
                        for (int k = 0; k < nr; k++)
                            for (int c = 0; c < J.c; c++) rtJ[c] += r[k] * J(k, c);
                        for (int k = 0; k < nr; k++)
                            for (int c = 0; c < J.c; c++) J(k, c) = sqrt_rho1 * (J(k, c) - alpha_sq_norm * r[k] * rtJ[c]);
                    }
                for (double& v : r) v *= residual_scaling;
            }
        }
        if (out) {
            ResEval& e = (*out)[ri];
            e.r = r;
            if (want_jac) {
                e.J.resize(nbk);
                for (int i = 0; i < nbk; i++) {
                    if (!raw[i]) continue;
                    const int local = p.blocks[rb.blocks[i]].local;
                    e.J[i] = local == Jg[i].c ? Jg[i] : Jg[i].block(0, 0, nr, local);  // times [I6; 0]
                }
            }
        }
    }
    return cost;
}

void scale_columns(const Program& g, const std::vector<double>& s, std::vector<ResEval>& ev) {
    Problem& p = *g.prob;
    for (size_t ri = 0; ri < ev.size(); ri++)
        for (size_t i = 0; i < ev[ri].J.size(); i++) {
            Mat& J = ev[ri].J[i];
            if (J.r == 0) continue;
            const int off = p.blocks[p.residuals[ri].blocks[i]].offset;
            for (int k = 0; k < J.r; k++)
                for (int c = 0; c < J.c; c++) J(k, c) *= s[off + c];
        }
}

std::vector<double> col_sq_norm(const Program& g, const std::vector<ResEval>& ev) {
    Problem& p = *g.prob;
    std::vector<double> n(g.ncols, 0.0);
    for (size_t ri = 0; ri < ev.size(); ri++)
        for (size_t i = 0; i < ev[ri].J.size(); i++) {
            const Mat& J = ev[ri].J[i];
            if (J.r == 0) continue;
            const int off = p.blocks[p.residuals[ri].blocks[i]].offset;
            for (int k = 0; k < J.r; k++)
                for (int c = 0; c < J.c; c++) n[off + c] += J(k, c) * J(k, c);
        }
    return n;
}

std::vector<double> JT_r(const Program& g, const std::vector<ResEval>& ev) {
    Problem& p = *g.prob;
    std::vector<double> v(g.ncols, 0.0);
    for (size_t ri = 0; ri < ev.size(); ri++)
        for (size_t i = 0; i < ev[ri].J.size(); i++) {
            const Mat& J = ev[ri].J[i];
            if (J.r == 0) continue;
            const int off = p.blocks[p.residuals[ri].blocks[i]].offset;
            for (int k = 0; k < J.r; k++)
                for (int c = 0; c < J.c; c++) v[off + c] += J(k, c) * ev[ri].r[k];
        }
    return v;
}

// per-residual-block J * v
std::vector<std::vector<double>> J_times(const Program& g, const std::vector<ResEval>& ev, const std::vector<double>& v) {
    Problem& p = *g.prob;
    std::vector<std::vector<double>> out(ev.size());
    for (size_t ri = 0; ri < ev.size(); ri++) {
        out[ri].assign(ev[ri].r.size(), 0.0);
        for (size_t i = 0; i < ev[ri].J.size(); i++) {
            const Mat& J = ev[ri].J[i];
            if (J.r == 0) continue;
            const int off = p.blocks[p.residuals[ri].blocks[i]].offset;
            for (int k = 0; k < J.r; k++)
                for (int c = 0; c < J.c; c++) out[ri][k] += J(k, c) * v[off + c];
        }
    }
    return out;
}

// DENSE_SCHUR: solves (J^T J + diag(D)^2) y = J^T r eliminating the landmark columns first.
bool dense_schur_solve(const Program& g, const std::vector<ResEval>& ev, const std::vector<double>& D,
                       const std::vector<double>& rhs, std::vector<double>& y) {
    Problem& p = *g.prob;
    const int np = g.n_other, nl = g.n_land;
    Mat Hpp(np, np);
    Mat Hlp(nl, np);
    std::vector<double> Hll(nl, 0.0);
    for (size_t ri = 0; ri < ev.size(); ri++) {
        const std::vector<int>& bl = p.residuals[ri].blocks;
        for (size_t i = 0; i < ev[ri].J.size(); i++) {
            const Mat& Ji = ev[ri].J[i];
            if (Ji.r == 0) continue;
            const int oi = p.blocks[bl[i]].offset;
            for (size_t j = 0; j < ev[ri].J.size(); j++) {
                const Mat& Jj = ev[ri].J[j];
                if (Jj.r == 0) continue;
                const int oj = p.blocks[bl[j]].offset;
                if (oi < np && oj < np) {
                    for (int a = 0; a < Ji.c; a++)
                        for (int c = 0; c < Jj.c; c++) {
                            double s = 0;
                            for (int k = 0; k < Ji.r; k++) s += Ji(k, a) * Jj(k, c);
                            Hpp(oi + a, oj + c) += s;
                        }
                } else if (oi >= np && oj < np) {
                    for (int c = 0; c < Jj.c; c++) {
                        double s = 0;
                        for (int k = 0; k < Ji.r; k++) s += Ji(k, 0) * Jj(k, c);
                        Hlp(oi - np, oj + c) += s;
                    }
                } else if (oi >= np && oj == oi) {
                    double s = 0;
                    for (int k = 0; k < Ji.r; k++) s += Ji(k, 0) * Ji(k, 0);
                    Hll[oi - np] += s;
                }
            }
        }
    }
    for (int i = 0; i < np; i++) Hpp(i, i) += D[i] * D[i];
    for (int l = 0; l < nl; l++) Hll[l] += D[np + l] * D[np + l];
    std::vector<double> bp(rhs.begin(), rhs.begin() + np);
    for (int l = 0; l < nl; l++) {
        if (!(Hll[l] > 0)) return false;
        const double inv = 1.0 / Hll[l];
        std::vector<int> nz;
        for (int c = 0; c < np; c++)
            if (Hlp(l, c) != 0) nz.push_back(c);
        for (int a : nz) {
            const double wa = Hlp(l, a) * inv;
            for (int c : nz) Hpp(a, c) -= wa * Hlp(l, c);
            bp[a] -= wa * rhs[np + l];
        }
    }
    Mat L;
    if (!cholesky(Hpp, L)) return false;
    y.assign(g.ncols, 0.0);
    chol_solve(L, bp.data(), y.data());
    for (int l = 0; l < nl; l++) {
        double s = rhs[np + l];
        for (int c = 0; c < np; c++) s -= Hlp(l, c) * y[c];
        y[np + l] = s / Hll[l];
    }
    for (double v : y)
        if (!std::isfinite(v)) return false;
    return true;
}

void plus(const Program& g, const std::vector<std::vector<double>>& x, const std::vector<double>& delta,
          std::vector<std::vector<double>>& out) {
    Problem& p = *g.prob;
    out = x;
    for (int b : g.active) {
        const Problem::Block& blk = p.blocks[b];
        if (blk.is_pose)
            pose_plus(x[b].data(), &delta[blk.offset], out[b].data());
        else
            for (int k = 0; k < blk.size; k++) out[b][k] = x[b][k] + delta[blk.offset + k];
    }
}

double norm_active(const Program& g, const std::vector<std::vector<double>>& x, const std::vector<std::vector<double>>* y) {
    double s = 0;
    for (int b : g.active)
        for (size_t k = 0; k < x[b].size(); k++) {
            const double d = y ? x[b][k] - (*y)[b][k] : x[b][k];
            s += d * d;
        }
    return std::sqrt(s);
}

double dot(const std::vector<double>& a, const std::vector<double>& b) {
    double s = 0;
    for (size_t i = 0; i < a.size(); i++) s += a[i] * b[i];
    return s;
}

}  // namespace

int ProbeColumns(Problem& problem) { return reduce_program(problem).ncols; }

int ProbeResiduals(Problem& problem) {
    int n = 0;
    for (auto& r : problem.residuals) n += r.cf->num_residuals;
    return n;
}

void ProbeEvaluate(Problem& problem, const double* delta, double* out) {
    Program g = reduce_program(problem);
    const int nb = (int)problem.blocks.size();
    std::vector<std::vector<double>> x(nb), xn;
    for (int b = 0; b < nb; b++) x[b].assign(problem.blocks[b].ptr, problem.blocks[b].ptr + problem.blocks[b].size);
    plus(g, x, std::vector<double>(delta, delta + g.ncols), xn);
    std::vector<const double*> params;
    int off = 0;
    for (auto& rb : problem.residuals) {
        const int nr = rb.cf->num_residuals;
        params.resize(rb.blocks.size());
        for (size_t i = 0; i < rb.blocks.size(); i++) params[i] = xn[rb.blocks[i]].data();
        rb.cf->Evaluate(params.data(), out + off, nullptr);
        if (rb.loss) {
            double sq = 0;
            for (int k = 0; k < nr; k++) sq += out[off + k] * out[off + k];
            if (sq > 0) {
                double rho[3];
                rb.loss->Evaluate(sq, rho);
                const double sc = std::sqrt(rho[0] / sq);
                for (int k = 0; k < nr; k++) out[off + k] *= sc;
            }
        }
        off += nr;
    }
}

SolveSummary Solve(Problem& problem, int max_num_iterations) {
    SolveSummary sum;
    Program g = reduce_program(problem);
    const int nb = (int)problem.blocks.size(), n = g.ncols;
    std::vector<std::vector<double>> x(nb), x_cand;
    for (int b = 0; b < nb; b++) x[b].assign(problem.blocks[b].ptr, problem.blocks[b].ptr + problem.blocks[b].size);
    auto write_back = [&]() {
        for (int b : g.active) std::memcpy(problem.blocks[b].ptr, x[b].data(), x[b].size() * sizeof(double));
    };
    if (n == 0) return sum;
    // --- iteration zero
    std::vector<ResEval> ev;
    double x_cost = evaluate(g, x, true, &ev);
    sum.initial_cost = sum.final_cost = x_cost;
    std::vector<double> scale(n);
    {
        const std::vector<double> cn = col_sq_norm(g, ev);
        for (int i = 0; i < n; i++) scale[i] = 1.0 / (1.0 + std::sqrt(cn[i]));  // Jacobi scaling, fixed for the solve
    }
    {
        const std::vector<double> grad = JT_r(g, ev);  // unscaled gradient
        double gmax = 0;
        for (double v : grad) gmax = std::max(gmax, std::fabs(v));
        if (gmax <= 1e-10) {
            sum.termination = 3;
            return sum;
        }
    }
    scale_columns(g, scale, ev);
    double x_norm = norm_active(g, x, nullptr);
    // DoglegStrategy state
    double radius = 1e4, mu = 1e-8, dogleg_step_norm = 0, alpha = 0;
    const double min_mu = 1e-8, max_mu = 1.0, mu_increase = 10.0;
    bool reuse = false;
    std::vector<double> diagonal(n), gradient(n), gauss_newton(n), step(n), delta(n);
    int num_consecutive_invalid = 0;
    for (int iteration = 1; iteration <= max_num_iterations; iteration++) {
        sum.iterations = iteration;
        bool linear_ok = true;
        if (!reuse) {
            reuse = true;
            const std::vector<double> cn = col_sq_norm(g, ev);
            for (int i = 0; i < n; i++) diagonal[i] = std::sqrt(std::min(std::max(cn[i], 1e-6), 1e32));
            gradient = JT_r(g, ev);
            for (int i = 0; i < n; i++) gradient[i] /= diagonal[i];
            {
                std::vector<double> sg(n);
                for (int i = 0; i < n; i++) sg[i] = gradient[i] / diagonal[i];
                const auto Jg = J_times(g, ev, sg);
                double jg2 = 0;
                for (auto& v : Jg)
                    for (double e : v) jg2 += e * e;
                alpha = dot(gradient, gradient) / jg2;
            }
            const std::vector<double> rhs = JT_r(g, ev);
            linear_ok = false;
            while (mu < max_mu) {
                std::vector<double> lm(n), y;
                for (int i = 0; i < n; i++) lm[i] = diagonal[i] * std::sqrt(mu);
                if (dense_schur_solve(g, ev, lm, rhs, y)) {
                    for (int i = 0; i < n; i++) gauss_newton[i] = -diagonal[i] * y[i];
                    linear_ok = true;
                    break;
                }
                mu *= mu_increase;
                sum.linear_solver_retries++;
            }
        }
        bool step_valid = false;
        double model_cost_change = 0;
        if (linear_ok) {
            // ComputeTraditionalDoglegStep
            const double gradient_norm = std::sqrt(dot(gradient, gradient));
            const double gn_norm = std::sqrt(dot(gauss_newton, gauss_newton));
            if (gn_norm <= radius) {
                step = gauss_newton;
                dogleg_step_norm = gn_norm;
            } else if (gradient_norm * alpha >= radius) {
                for (int i = 0; i < n; i++) step[i] = -(radius / gradient_norm) * gradient[i];
                dogleg_step_norm = radius;
            } else {
                const double b_dot_a = -alpha * dot(gradient, gauss_newton);
                const double a_squared_norm = std::pow(alpha * gradient_norm, 2.0);
                const double b_minus_a_squared_norm = a_squared_norm - 2 * b_dot_a + std::pow(gn_norm, 2);
                const double c = b_dot_a - a_squared_norm;
                const double d = std::sqrt(c * c + b_minus_a_squared_norm * (std::pow(radius, 2.0) - a_squared_norm));
                const double beta = (c <= 0) ? (d - c) / b_minus_a_squared_norm : (radius * radius - a_squared_norm) / (d + c);
                for (int i = 0; i < n; i++) step[i] = (-alpha * (1.0 - beta)) * gradient[i] + beta * gauss_newton[i];
                dogleg_step_norm = std::sqrt(dot(step, step));
            }
            for (int i = 0; i < n; i++) step[i] /= diagonal[i];
            const auto mr = J_times(g, ev, step);
            for (size_t ri = 0; ri < mr.size(); ri++)
                for (size_t k = 0; k < mr[ri].size(); k++) model_cost_change -= mr[ri][k] * (ev[ri].r[k] + mr[ri][k] / 2.0);
            step_valid = model_cost_change > 0.0;
        }
        if (!step_valid) {  // HandleInvalidStep
            if (++num_consecutive_invalid >= 5) {
                sum.termination = 4;
                break;
            }
            mu *= mu_increase;  // StepIsInvalid
            reuse = false;
            continue;
        }
        num_consecutive_invalid = 0;
        for (int i = 0; i < n; i++) delta[i] = step[i] * scale[i];
        plus(g, x, delta, x_cand);
        const double cand_cost = evaluate(g, x_cand, false, nullptr);
        const double step_norm = norm_active(g, x, &x_cand);
        if (step_norm <= 1e-8 * (x_norm + 1e-8)) {  // parameter_tolerance
            sum.termination = 1;
            break;
        }
        const double cost_change = x_cost - cand_cost;
        if (std::fabs(cost_change) <= 1e-6 * x_cost) {  // function_tolerance
            sum.termination = 2;
            break;
        }
        const double relative_decrease = cost_change / model_cost_change;
        if (relative_decrease > 1e-3) {  // min_relative_decrease: successful step
            x = x_cand;
            x_norm = norm_active(g, x, nullptr);
            x_cost = evaluate(g, x, true, &ev);
            const std::vector<double> grad = JT_r(g, ev);
            scale_columns(g, scale, ev);
            sum.successful_steps++;
            // DoglegStrategy::StepAccepted
            if (relative_decrease < 0.25) radius *= 0.5;
            if (relative_decrease > 0.75) radius = std::max(radius, 3.0 * dogleg_step_norm);
            mu = std::max(min_mu, 2.0 * mu / mu_increase);
            reuse = false;
            double gmax = 0;
            for (double v : grad) gmax = std::max(gmax, std::fabs(v));
            if (gmax <= 1e-10) {
                sum.termination = 3;
                break;
            }
        } else {  // StepRejected
            radius *= 0.5;
            reuse = true;
        }
    }
    sum.final_cost = x_cost;
    write_back();
    return sum;
}

}  // namespace orc

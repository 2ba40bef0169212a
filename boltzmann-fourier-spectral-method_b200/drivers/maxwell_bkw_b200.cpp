// maxwell_bkw_b200 -- BKW accuracy/timing driver for BoltzmannOperator<B200_Backend>.
//
// Caller side of the drop-in boundary: same experiment, constants and printed output as the
// reference drivers maxwell_bkw_fftw.cpp:23-174 / maxwell_bkw_cuda.cu:24-187 (Maxwell molecules,
// BKW solution at t = 6.5, L1/L2/Linf error of Q against the analytic time derivative), with
//   --Nv N --Ns S -t T          as in the reference
//   --Nr R                      Gauss-Legendre point count (the reference ties it to Nv)
//   --design-dir DIR            directory with the ssTTT.NNN.txt node files ($BFSM_DESIGN_DIR)
//   --backend b200|fftw|both    `fftw`/`both` only when built with -DBFSM_HAVE_REFERENCE_HEADERS
//                               inside the build container (links the reference CPU operator)
//   --gpus G                    pair-shard the evaluation over G GPUs of this host: one operator per
//                               GPU, ncclCommInitAll, one all-reduce per evaluation
//   --t0 A --tfinal B --dt H    BASELINE config 3: integrate df/dt = Q(f,f) from the exact BKW solution
//                               at t0 to tfinal with classical RK4 (all stages on the device), report the
//                               error against the exact solution at tfinal and the conservation defects.
//                               The reference has no time integrator (its drivers evaluate Q once);
//                               with --backend both the SAME loop is driven by the FFTW backend and the
//                               two final states are compared.
// Differences on purpose: argument errors are fatal (the reference prints and continues with
// uninitialised values, maxwell_bkw_fftw.cpp:50-51); Linf is a true max (the reference's
// OpenMP `reduction(+)` around a max is only right with one thread, :148-156).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <iomanip>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

#ifdef BFSM_HAVE_REFERENCE_HEADERS
#include "Collisions/FFTWBoltzmannOperator.hpp"
#include "Utilities/statistics.hpp"
#endif
#include "B200BoltzmannOperator.hpp"

namespace {

double now_s()
{
    using clk = std::chrono::steady_clock;
    return std::chrono::duration<double>(clk::now().time_since_epoch()).count();
}

void stats(const std::string &name, const std::vector<double> &t)
{
#ifdef BFSM_HAVE_REFERENCE_HEADERS
    print_stats_summary(name, t);
#else
    double mn = t[0], mx = t[0], sum = 0;
    for (double v : t) {
        mn = std::min(mn, v);
        mx = std::max(mx, v);
        sum += v;
    }
    const double mean = sum / t.size();
    double ss = 0;
    for (double v : t) ss += (v - mean) * (v - mean);
    const double sd = std::sqrt(ss / std::max<size_t>(1, t.size() - 1));
    std::cout << "\nRun statistics for " << name << "\n";
    std::cout << "Total number of samples taken: " << t.size() << "\n";
    std::cout << std::scientific << std::setprecision(8) << "Mean runtime (s): " << mean << "\n";
    std::cout << "Min runtime (s): " << mn << "\n";
    std::cout << "Max runtime (s): " << mx << "\n";
    std::cout << "stdev: " << sd << "\n\n";
#endif
}

void report_errors(const char *what, const std::vector<double> &Q, const std::vector<double> &Q_exact, double dv)
{
    double l1 = 0, l2 = 0, linf = 0;
    for (size_t i = 0; i < Q.size(); ++i) {
        const double d = std::abs(Q[i] - Q_exact[i]);
        l1 += d;
        l2 += d * d;
        linf = std::max(linf, d);
    }
    l1 *= dv * dv * dv;
    l2 = std::sqrt(l2 * dv * dv * dv);
    std::cout << std::scientific << std::setprecision(8);
    std::cout << what << ":\n";
    std::cout << "L1 error: " << l1 << "\n";
    std::cout << "L2 error: " << l2 << "\n";
    std::cout << "Linf error: " << linf << "\n\n";
}

[[noreturn]] void usage(const char *msg)
{
    std::cerr << "error: " << msg << "\n"
              << "usage: maxwell_bkw_b200 [--Nv N] [--Nr R] [--Ns S] [-t trials] [--backend b200|fftw|both]"
                 " [--design-dir DIR] [--device D] [--gpus G] [--t0 A --tfinal B --dt H]\n";
    std::exit(2);
}

void ck(int rc)
{
    if (rc != BFSM_OK) throw std::runtime_error(bfsm_last_error());
}

// exact BKW solution and its time derivative on the grid (maxwell_bkw_fftw.cpp:74-99)
void bkw(int Nv, double L, double t, std::vector<double> &f, std::vector<double> *dfdt)
{
    const double dv = 2 * L / Nv;
    const double K = 1 - std::exp(-t / 6), dK = std::exp(-t / 6) / 6;
    const double norm = 1 / (2 * std::pow(2 * pi * K, 1.5));
    std::vector<double> v(Nv);
    for (int i = 0; i < Nv; ++i) v[i] = -L + dv / 2 + i * dv;
    for (int i = 0; i < Nv; ++i)
        for (int j = 0; j < Nv; ++j)
            for (int k = 0; k < Nv; ++k) {
                const size_t idx = ((size_t)i * Nv + j) * Nv + k;
                const double r_sq = v[i] * v[i] + v[j] * v[j] + v[k] * v[k];
                const double g = std::exp(-r_sq / (2 * K));
                f[idx] = g * ((5 * K - 3) / K + (1 - K) / (K * K) * r_sq) * norm;
                if (dfdt) {
                    double q = (-3 / (2 * K) + r_sq / (2 * K * K)) * f[idx];
                    q += norm * g * (3 / (K * K) + (K - 2) / (K * K * K) * r_sq);
                    (*dfdt)[idx] = q * dK;
                }
            }
}

// Classical RK4 on buffers the backend owns: collide(Q, f), axpby(out, a, x, b, y).
template <class Buf>
int rk4(const std::function<void(Buf, Buf)> &collide, const std::function<void(Buf, double, Buf, double, Buf)> &axpby,
        Buf f, Buf k0, Buf k1, Buf k2, Buf k3, Buf tmp, double t0, double t_final, double dt)
{
    const int n_steps = std::max(1, (int)std::lround((t_final - t0) / dt));
    const double h = (t_final - t0) / n_steps;
    for (int s = 0; s < n_steps; ++s) {
        collide(k0, f);
        axpby(tmp, 1.0, f, 0.5 * h, k0);
        collide(k1, tmp);
        axpby(tmp, 1.0, f, 0.5 * h, k1);
        collide(k2, tmp);
        axpby(tmp, 1.0, f, h, k2);
        collide(k3, tmp);
        axpby(f, 1.0, f, h / 6, k0);
        axpby(f, 1.0, f, h / 3, k1);
        axpby(f, 1.0, f, h / 3, k2);
        axpby(f, 1.0, f, h / 6, k3);
    }
    return n_steps;
}

} // namespace

int main(int argc, char **argv)
{
    int Nv = 32, Ns = 12, Nr = -1, trials = 1, device = 0, gpus = 1;
    double t0 = 6.5, t_final = -1, dt = 0.1;
    std::string backend = "b200", design_dir;
    for (int a = 1; a < argc; ++a) {
        const std::string k = argv[a];
        auto need = [&](const char *name) -> const char * {
            if (a + 1 >= argc) usage((std::string("missing value for ") + name).c_str());
            return argv[++a];
        };
        if (k == "--Nv") Nv = std::atoi(need("--Nv"));
        else if (k == "--Ns") Ns = std::atoi(need("--Ns"));
        else if (k == "--Nr") Nr = std::atoi(need("--Nr"));
        else if (k == "-t" || k == "--trials") trials = std::atoi(need("--trials"));
        else if (k == "--backend") backend = need("--backend");
        else if (k == "--design-dir") design_dir = need("--design-dir");
        else if (k == "--device") device = std::atoi(need("--device"));
        else if (k == "--gpus") gpus = std::atoi(need("--gpus"));
        else if (k == "--t0") t0 = std::atof(need("--t0"));
        else if (k == "--tfinal") t_final = std::atof(need("--tfinal"));
        else if (k == "--dt") dt = std::atof(need("--dt"));
        else usage(("unknown argument " + k).c_str());
    }
    if (Nr <= 0) Nr = Nv; // reference behaviour: maxwell_bkw_fftw.cpp:102
    if (Nv <= 0 || Ns <= 0 || trials <= 0 || gpus <= 0) usage("Nv, Ns, trials and gpus must be positive");
    const bool integrate = t_final > t0;
    if (integrate && !(dt > 0)) usage("--dt must be positive");
    if (integrate && t0 <= 6 * std::log(2.5)) usage("--t0 must exceed 6 ln(5/2) ~ 5.498 (BKW positivity)");
    if (!design_dir.empty()) setenv("BFSM_DESIGN_DIR", design_dir.c_str(), 1);
#ifdef BFSM_HAVE_REFERENCE_HEADERS
    if (!design_dir.empty()) setenv("BFSM_REF_DESIGN_DIR", design_dir.c_str(), 1);
#else
    if (backend != "b200") usage("this build has only the b200 backend");
#endif

    std::cout << "\nRun arguments:\n";
    std::cout << "Nv = " << Nv << "\nNr = " << Nr << "\nNs = " << Ns << "\ntrials = " << trials
              << "\nbackend = " << backend << "\ngpus = " << gpus << "\n";
    if (integrate) std::cout << "t0 = " << t0 << "\ntfinal = " << t_final << "\ndt = " << dt << "\n";

    // constants of the experiment (maxwell_bkw_fftw.cpp:54-60, 74-76)
    const double gamma = 0;
    const double b_gamma = 1 / (4 * pi);
    const double S = 5, R = 2 * S;
    const double L = ((3 + std::sqrt(2.0)) / 2) * S;
    const double dv = 2 * L / Nv;
    const double t = integrate ? t0 : 6.5;

    const size_t N = (size_t)Nv * Nv * Nv;
    std::vector<double> f_bkw(N), Q_bkw(N), Q(N), f_exact_final(N), f_b200_final;
    bkw(Nv, L, t, f_bkw, &Q_bkw);
    if (integrate) bkw(Nv, L, t_final, f_exact_final, nullptr);

    try {
        auto gl = std::make_shared<GaussLegendreQuadrature>(Nr, 0, R);
        auto sph = std::make_shared<SphericalDesign>(Ns);

        if (backend == "b200" || backend == "both") {
            typedef BoltzmannOperator<B200_Backend> Op;
            std::vector<std::unique_ptr<Op>> ops;
            std::vector<bfsm_comm *> comms(gpus, nullptr);
            std::vector<int> devs(gpus);
            for (int g = 0; g < gpus; ++g) devs[g] = device + g;
            if (gpus > 1) ck(bfsm_comm_init_all(comms.data(), gpus, devs.data()));
            double ti = now_s();
            for (int g = 0; g < gpus; ++g) {
                ops.emplace_back(new Op(gl, sph, Nv, Nv, Nv, gamma, b_gamma, L));
                ops[g]->setDevice(devs[g]);
                if (gpus > 1) {
                    ops[g]->setShard(g, gpus);
                    ops[g]->setCommunicator(comms[g]);
                }
                ops[g]->initialize();
            }
            std::cout << "Initialization time (s): " << now_s() - ti << " seconds\n";

            // device buffers, host->device copy outside the timed loop (maxwell_bkw_cuda.cu:119-126);
            // per GPU: f, k0..k3, tmp
            const size_t bytes = N * sizeof(double);
            std::vector<std::vector<double *>> buf(gpus, std::vector<double *>(6, nullptr));
            for (int g = 0; g < gpus; ++g) {
                for (auto &b : buf[g]) ck(bfsm_device_malloc(devs[g], (void **)&b, bytes));
                ck(bfsm_copy_to_device(devs[g], buf[g][0], f_bkw.data(), bytes));
            }
            // one evaluation on every GPU: buffers a (out) and b (in) by index
            auto collide = [&](int a, int b) {
                if (gpus == 1) {
                    (*ops[0])(buf[0][a], buf[0][b]);
                } else {
                    std::vector<Op *> raw;
                    std::vector<double *> q;
                    std::vector<const double *> fin;
                    for (int g = 0; g < gpus; ++g) {
                        raw.push_back(ops[g].get());
                        q.push_back(buf[g][a]);
                        fin.push_back(buf[g][b]);
                    }
                    Op::computeCollisionGroup(raw, q, fin);
                }
            };
            std::vector<double> times;
            for (int trial = 0; trial < trials; ++trial) {
                ti = now_s();
                collide(1, 0);
                times.push_back(now_s() - ti);
            }
            stats(ops[0]->getBackendName(), times);
            ck(bfsm_copy_to_host(devs[0], Q.data(), buf[0][1], bytes));
            report_errors("Approximation errors", Q, Q_bkw, dv);

            // conservation defects of Q (mass, momentum, energy of the collision term vanish)
            double *m_dev = nullptr;
            double m[5];
            ck(bfsm_device_malloc(devs[0], (void **)&m_dev, sizeof m));
            ck(bfsm_moments(ops[0]->handle(), buf[0][1], 1, m_dev, nullptr));
            ck(bfsm_copy_to_host(devs[0], m, m_dev, sizeof m));
            std::cout << "Moments of Q (mass, momentum x y z, energy): " << m[0] << " " << m[1] << " " << m[2]
                      << " " << m[3] << " " << m[4] << "\n\n";

            if (integrate) {
                ti = now_s();
                const int steps = rk4<int>(
                    [&](int a, int b) { collide(a, b); },
                    [&](int out, double a, int x, double b, int y) {
                        for (int g = 0; g < gpus; ++g)
                            ck(bfsm_vec_axpby(devs[g], buf[g][out], a, buf[g][x], b, buf[g][y], N, nullptr));
                    },
                    0, 1, 2, 3, 4, 5, t0, t_final, dt);
                f_b200_final.resize(N);
                ck(bfsm_copy_to_host(devs[0], f_b200_final.data(), buf[0][0], bytes));
                std::cout << "RK4: " << steps << " steps, " << 4 * steps << " evaluations of Q in "
                          << now_s() - ti << " s\n";
                report_errors("Error of f(tfinal) against the exact BKW solution", f_b200_final, f_exact_final, dv);
                ck(bfsm_moments(ops[0]->handle(), buf[0][0], 1, m_dev, nullptr));
                ck(bfsm_copy_to_host(devs[0], m, m_dev, sizeof m));
                std::cout << "Moments of f(tfinal) (mass, momentum x y z, energy): " << m[0] << " " << m[1] << " "
                          << m[2] << " " << m[3] << " " << m[4] << "   (exact: 1 0 0 0 1.5)\n\n";
            }
            bfsm_device_free(devs[0], m_dev);
            for (int g = 0; g < gpus; ++g)
                for (auto b : buf[g]) bfsm_device_free(devs[g], b);
            ops.clear();
            for (auto c : comms) bfsm_comm_destroy(c);
        }
#ifdef BFSM_HAVE_REFERENCE_HEADERS
        if (backend == "fftw" || backend == "both") {
            std::vector<double> Q_ref(N);
            BoltzmannOperator<FFTW_Backend> op(gl, sph, Nv, Nv, Nv, gamma, b_gamma, L);
            double ti = now_s();
            op.initialize();
            std::cout << "Initialization time (s): " << now_s() - ti << " seconds\n";
            std::vector<double> times;
            for (int trial = 0; trial < trials; ++trial) {
                ti = now_s();
                op(Q_ref.data(), f_bkw.data());
                times.push_back(now_s() - ti);
            }
            stats(op.getBackendName(), times);
            report_errors("Approximation errors", Q_ref, Q_bkw, dv);
            if (backend == "both") {
                double num = 0, den = 0;
                for (size_t i = 0; i < N; ++i) {
                    num = std::max(num, std::abs(Q[i] - Q_ref[i]));
                    den = std::max(den, std::abs(Q_ref[i]));
                }
                std::cout << "B200 vs FFTW: max|dQ|/max|Q| = " << num / den << "\n";
            }
            if (integrate) {
                // the same RK4 loop on host buffers, driven by the reference CPU operator
                std::vector<std::vector<double>> hb(6, std::vector<double>(N));
                hb[0] = f_bkw;
                ti = now_s();
                const int steps = rk4<int>(
                    [&](int a, int b) { op(hb[a].data(), hb[b].data()); },
                    [&](int out, double a, int x, double b, int y) {
                        for (size_t i = 0; i < N; ++i) hb[out][i] = a * hb[x][i] + b * hb[y][i];
                    },
                    0, 1, 2, 3, 4, 5, t0, t_final, dt);
                std::cout << "RK4 (FFTW backend): " << steps << " steps in " << now_s() - ti << " s\n";
                report_errors("Error of f(tfinal) against the exact BKW solution (FFTW backend)", hb[0], f_exact_final, dv);
                if (backend == "both" && !f_b200_final.empty()) {
                    double num = 0, den = 0;
                    for (size_t i = 0; i < N; ++i) {
                        num = std::max(num, std::abs(f_b200_final[i] - hb[0][i]));
                        den = std::max(den, std::abs(hb[0][i]));
                    }
                    std::cout << "B200 vs FFTW after integration: max|df|/max|f| = " << num / den << "\n";
                }
            }
        }
#endif
    } catch (const std::exception &e) {
        std::cerr << "error: " << e.what() << "\n";
        return 1;
    }
    return 0;
}

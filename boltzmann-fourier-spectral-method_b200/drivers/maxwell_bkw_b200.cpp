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
// Differences on purpose: argument errors are fatal (the reference prints and continues with
// uninitialised values, maxwell_bkw_fftw.cpp:50-51); Linf is a true max (the reference's
// OpenMP `reduction(+)` around a max is only right with one thread, :148-156).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
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

void report_errors(const std::vector<double> &Q, const std::vector<double> &Q_exact, double dv)
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
    std::cout << "Approximation errors:\n";
    std::cout << "L1 error: " << l1 << "\n";
    std::cout << "L2 error: " << l2 << "\n";
    std::cout << "Linf error: " << linf << "\n\n";
}

[[noreturn]] void usage(const char *msg)
{
    std::cerr << "error: " << msg << "\n"
              << "usage: maxwell_bkw_b200 [--Nv N] [--Nr R] [--Ns S] [-t trials] [--backend b200|fftw|both]"
                 " [--design-dir DIR] [--device D]\n";
    std::exit(2);
}

} // namespace

int main(int argc, char **argv)
{
    int Nv = 32, Ns = 12, Nr = -1, trials = 1, device = 0;
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
        else usage(("unknown argument " + k).c_str());
    }
    if (Nr <= 0) Nr = Nv; // reference behaviour: maxwell_bkw_fftw.cpp:102
    if (Nv <= 0 || Ns <= 0 || trials <= 0) usage("Nv, Ns and trials must be positive");
    if (!design_dir.empty()) setenv("BFSM_DESIGN_DIR", design_dir.c_str(), 1);
#ifdef BFSM_HAVE_REFERENCE_HEADERS
    if (!design_dir.empty()) setenv("BFSM_REF_DESIGN_DIR", design_dir.c_str(), 1);
#else
    if (backend != "b200") usage("this build has only the b200 backend");
#endif

    std::cout << "\nRun arguments:\n";
    std::cout << "Nv = " << Nv << "\nNr = " << Nr << "\nNs = " << Ns << "\ntrials = " << trials
              << "\nbackend = " << backend << "\n";

    // constants of the experiment (maxwell_bkw_fftw.cpp:54-60, 74-76)
    const double gamma = 0;
    const double b_gamma = 1 / (4 * pi);
    const double S = 5, R = 2 * S;
    const double L = ((3 + std::sqrt(2.0)) / 2) * S;
    const double dv = 2 * L / Nv;
    const double t = 6.5;
    const double K = 1 - std::exp(-t / 6);
    const double dK = std::exp(-t / 6) / 6;

    const size_t N = (size_t)Nv * Nv * Nv;
    std::vector<double> v(Nv), f_bkw(N), Q_bkw(N), Q(N);
    for (int i = 0; i < Nv; ++i) v[i] = -L + dv / 2 + i * dv;
    const double norm = 1 / (2 * std::pow(2 * pi * K, 1.5));
    for (int i = 0; i < Nv; ++i)
        for (int j = 0; j < Nv; ++j)
            for (int k = 0; k < Nv; ++k) {
                const size_t idx = ((size_t)i * Nv + j) * Nv + k;
                const double r_sq = v[i] * v[i] + v[j] * v[j] + v[k] * v[k];
                const double g = std::exp(-r_sq / (2 * K));
                f_bkw[idx] = g * ((5 * K - 3) / K + (1 - K) / (K * K) * r_sq) * norm;
                double q = (-3 / (2 * K) + r_sq / (2 * K * K)) * f_bkw[idx];
                q += norm * g * (3 / (K * K) + (K - 2) / (K * K * K) * r_sq);
                Q_bkw[idx] = q * dK;
            }

    try {
        auto gl = std::make_shared<GaussLegendreQuadrature>(Nr, 0, R);
        auto sph = std::make_shared<SphericalDesign>(Ns);

        if (backend == "b200" || backend == "both") {
            BoltzmannOperator<B200_Backend> op(gl, sph, Nv, Nv, Nv, gamma, b_gamma, L);
            op.setDevice(device);
            double t0 = now_s();
            op.initialize();
            std::cout << "Initialization time (s): " << now_s() - t0 << " seconds\n";

            // device buffers, host->device copy outside the timed loop (maxwell_bkw_cuda.cu:119-126)
            double *f_dev = nullptr, *Q_dev = nullptr;
            if (bfsm_device_malloc(device, (void **)&f_dev, N * sizeof(double)) ||
                bfsm_device_malloc(device, (void **)&Q_dev, N * sizeof(double)) ||
                bfsm_copy_to_device(device, f_dev, f_bkw.data(), N * sizeof(double)))
                throw std::runtime_error(bfsm_last_error());
            std::vector<double> times;
            for (int trial = 0; trial < trials; ++trial) {
                t0 = now_s();
                op(Q_dev, f_dev);
                times.push_back(now_s() - t0);
            }
            stats(op.getBackendName(), times);
            if (bfsm_copy_to_host(device, Q.data(), Q_dev, N * sizeof(double)))
                throw std::runtime_error(bfsm_last_error());
            report_errors(Q, Q_bkw, dv);
            bfsm_device_free(device, f_dev);
            bfsm_device_free(device, Q_dev);
        }
#ifdef BFSM_HAVE_REFERENCE_HEADERS
        if (backend == "fftw" || backend == "both") {
            std::vector<double> Q_ref(N);
            BoltzmannOperator<FFTW_Backend> op(gl, sph, Nv, Nv, Nv, gamma, b_gamma, L);
            double t0 = now_s();
            op.initialize();
            std::cout << "Initialization time (s): " << now_s() - t0 << " seconds\n";
            std::vector<double> times;
            for (int trial = 0; trial < trials; ++trial) {
                t0 = now_s();
                op(Q_ref.data(), f_bkw.data());
                times.push_back(now_s() - t0);
            }
            stats(op.getBackendName(), times);
            report_errors(Q_ref, Q_bkw, dv);
            if (backend == "both") {
                double num = 0, den = 0;
                for (size_t i = 0; i < N; ++i) {
                    num = std::max(num, std::abs(Q[i] - Q_ref[i]));
                    den = std::max(den, std::abs(Q_ref[i]));
                }
                std::cout << "B200 vs FFTW: max|dQ|/max|Q| = " << num / den << "\n";
            }
        }
#endif
    } catch (const std::exception &e) {
        std::cerr << "error: " << e.what() << "\n";
        return 1;
    }
    return 0;
}

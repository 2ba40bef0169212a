// bfsm_host.hpp -- self-contained host-side pieces for builds WITHOUT the reference tree.
//
// When B200BoltzmannOperator.hpp is compiled inside the reference repository it uses the
// reference's own AbstractCollisionOperator / GaussLegendreQuadrature / SphericalQuadrature
// (Collisions/AbstractCollisionOperator.hpp, Quadratures/*.hpp).  Outside of it (this repo's
// standalone driver, the GPU box) the equivalents below are used instead; they expose the same
// member names so the operator class and drivers compile unchanged against either set.
//
//   GaussLegendreQuadrature(n,a,b)  GSL-free: Newton iteration on P_n in long double, nodes
//                                   ascending as in Quadratures/GaussLegendre.hpp:10-24
//   SphericalDesign(N[,dir])        reads ssTTT.NNN.txt ("x y z" per line) from a directory
//                                   ($BFSM_DESIGN_DIR or the argument) instead of the absolute
//                                   paths hard-coded in Quadratures/SphericalDesign.cpp:13-21;
//                                   equal weights 4*pi/N (SphericalDesign.cpp:48)
#ifndef BFSM_HOST_HPP
#define BFSM_HOST_HPP

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#ifndef BFSM_HAVE_REFERENCE_HEADERS

constexpr double pi = 3.14159265358979323846;

// Interface every collision backend implements (same five members as the reference's base).
class AbstractCollisionOperator {
public:
    virtual ~AbstractCollisionOperator() = default;
    virtual void initialize() = 0;
    virtual std::string getBackendName() const = 0;
    virtual void computeCollision(double *Q, const double *f_in) = 0;
    virtual void operator()(double *Q, const double *f_in) = 0;
};

template <typename Backend> class BoltzmannOperator;

class GaussLegendreQuadrature {
public:
    GaussLegendreQuadrature(int n_points, double a, double b)
    {
        if (n_points <= 0) throw std::invalid_argument("Gauss-Legendre: n_points must be positive");
        const int n = n_points, m = (n + 1) / 2;
        std::vector<long double> xh(m), wh(m); // non-negative half, ascending
        const long double PI_L = 3.14159265358979323846264338327950288L;
        for (int k = 1; k <= m; ++k) {
            long double z = std::cos(PI_L * (k - 0.25L) / (n + 0.5L));
            long double p0, p1, pp = 1;
            auto eval = [&](long double x) {
                p0 = 1;
                p1 = x;
                for (int j = 2; j <= n; ++j) {
                    const long double p2 = ((2.0L * j - 1) * x * p1 - (j - 1.0L) * p0) / j;
                    p0 = p1;
                    p1 = p2;
                }
                pp = n * (x * p1 - p0) / (x * x - 1);
            };
            for (int it = 0; it < 100; ++it) {
                eval(z);
                const long double dz = p1 / pp;
                z -= dz;
                if (std::fabs(dz) < 1e-20L) break;
            }
            eval(z);
            if ((n & 1) && k == m) z = 0;
            xh[m - k] = z;
            wh[m - k] = 2 / ((1 - z * z) * pp * pp);
        }
        const double A = (b - a) / 2, B = (a + b) / 2;
        nodes.resize(n);
        weights.resize(n);
        for (int i = 0; i < n; ++i) {
            int k;
            double sgn;
            if (n & 1) {
                k = i - n / 2;
                sgn = k < 0 ? -1.0 : 1.0;
                k = k < 0 ? -k : k;
            } else if (i < n / 2) {
                k = n / 2 - 1 - i;
                sgn = -1.0;
            } else {
                k = i - n / 2;
                sgn = 1.0;
            }
            nodes[i] = B + sgn * A * (double)xh[k];
            weights[i] = A * (double)wh[k];
        }
    }
    const std::vector<double> &getWeights() const { return weights; }
    const std::vector<double> &getNodes() const { return nodes; }
    int getNumberOfPoints() const { return (int)weights.size(); }

protected:
    std::vector<double> weights, nodes;
};

class SphericalQuadrature {
public:
    virtual ~SphericalQuadrature() = default;
    const std::vector<double> &getWeights() const { return weights; }
    const std::vector<double> &getx() const { return x; }
    const std::vector<double> &gety() const { return y; }
    const std::vector<double> &getz() const { return z; }
    int getNumberOfPoints() const { return (int)weights.size(); }

protected:
    std::vector<double> weights, x, y, z;
};

class SphericalDesign : public SphericalQuadrature {
public:
    explicit SphericalDesign(int N, const std::string &design_dir = std::string())
    {
        if (N <= 0) throw std::invalid_argument("Number of points N must be a positive integer");
        static const int sizes[9] = {6, 12, 32, 48, 70, 94, 120, 156, 192};
        static const int degrees[9] = {3, 5, 7, 9, 11, 13, 15, 17, 19};
        int degree = -1;
        for (int i = 0; i < 9; ++i)
            if (sizes[i] == N) degree = degrees[i];
        if (degree < 0) throw std::invalid_argument("Invalid value of N");
        std::string dir = design_dir;
        if (dir.empty()) {
            const char *env = std::getenv("BFSM_DESIGN_DIR");
            dir = env ? env : ".";
        }
        char name[64];
        std::snprintf(name, sizeof name, "/ss%03d.%03d.txt", degree, N);
        const std::string filename = dir + name;
        std::ifstream file(filename);
        if (!file.is_open()) throw std::runtime_error("Could not open file " + filename);
        std::string line;
        while (std::getline(file, line)) {
            std::istringstream iss(line);
            double a, b, c;
            if (iss >> a >> b >> c) {
                x.push_back(a);
                y.push_back(b);
                z.push_back(c);
            }
        }
        if ((int)x.size() != N) throw std::runtime_error("design file has the wrong number of rows: " + filename);
        weights.assign(N, (4 * pi) / N);
    }
};

#endif // !BFSM_HAVE_REFERENCE_HEADERS
#endif // BFSM_HOST_HPP

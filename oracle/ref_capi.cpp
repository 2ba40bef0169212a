// oracle/ref_capi.cpp -- TEST INFRASTRUCTURE, not product code.
//
// extern "C" handle API around the UNMODIFIED reference CPU operator
// BoltzmannOperator<FFTW_Backend> (/root/reference/Collisions/FFTWBoltzmannOperator.{hpp,cpp})
// and its quadrature classes (/root/reference/Quadratures/*), compiled in place
// from /root/reference by oracle/Makefile against the FFTW/GSL stand-ins in
// oracle/shim/.  The resulting oracle/_ref/libbfsm_ref.so is what tests and
// bench.py call "the reference" (cpu_baseline.kind == "reference").
//
// Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs load it.
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <exception>

#include <dlfcn.h>
#include <omp.h>

#include "Collisions/FFTWBoltzmannOperator.hpp"

namespace {

struct RefHandle {
    std::shared_ptr<GaussLegendreQuadrature> gl;
    std::shared_ptr<SphericalQuadrature> sph;
    std::unique_ptr<BoltzmannOperator<FFTW_Backend>> op;
    int n;
};

thread_local std::string g_err;

// The patched copy of SphericalDesign.cpp reads $BFSM_REF_DESIGN_DIR; default it
// to the "designs" directory that sits next to this shared object.
void ensure_design_dir()
{
    if (std::getenv("BFSM_REF_DESIGN_DIR")) return;
    Dl_info info;
    if (dladdr((void *)&ensure_design_dir, &info) && info.dli_fname) {
        std::string p(info.dli_fname);
        size_t k = p.find_last_of('/');
        std::string dir = (k == std::string::npos ? std::string(".") : p.substr(0, k)) + "/designs";
        setenv("BFSM_REF_DESIGN_DIR", dir.c_str(), 0);
    }
}

} // namespace

extern "C" {

const char *bfsm_ref_last_error(void) { return g_err.c_str(); }

int bfsm_ref_max_threads(void) { return omp_get_max_threads(); }

void bfsm_ref_set_threads(int n)
{
    if (n > 0) omp_set_num_threads(n);
}

// Mirrors the construction sequence of maxwell_bkw_fftw.cpp:102-114, with the
// Gauss-Legendre point count decoupled from Nv.
void *bfsm_ref_create(int nvx, int nvy, int nvz, int n_gl, double a, double b, int n_sph,
                      double gamma, double b_gamma, double L)
{
    try {
        ensure_design_dir();
        std::unique_ptr<RefHandle> h(new RefHandle);
        h->gl = std::make_shared<GaussLegendreQuadrature>(n_gl, a, b);
        h->sph = std::make_shared<SphericalDesign>(n_sph);
        if (h->sph->getx().size() != (size_t)n_sph) {
            g_err = "spherical design file has the wrong number of rows";
            return nullptr;
        }
        h->op.reset(new BoltzmannOperator<FFTW_Backend>(h->gl, h->sph, nvx, nvy, nvz, gamma,
                                                        b_gamma, L));
        h->op->initialize();
        h->n = nvx * nvy * nvz;
        return h.release();
    } catch (const std::exception &e) {
        g_err = e.what();
        return nullptr;
    }
}

int bfsm_ref_apply(void *handle, const double *f_in, double *Q)
{
    RefHandle *h = static_cast<RefHandle *>(handle);
    if (!h) return 1;
    (*h->op)(Q, f_in);
    return 0;
}

// Returns wall seconds (omp_get_wtime, as the reference driver does,
// maxwell_bkw_fftw.cpp:133-140) of ONE evaluation.
double bfsm_ref_apply_timed(void *handle, const double *f_in, double *Q)
{
    RefHandle *h = static_cast<RefHandle *>(handle);
    if (!h) return -1.0;
    double t0 = omp_get_wtime();
    (*h->op)(Q, f_in);
    return omp_get_wtime() - t0;
}

int bfsm_ref_quadrature(void *handle, double *gl_nodes, double *gl_wts, double *sx, double *sy,
                        double *sz, double *sw)
{
    RefHandle *h = static_cast<RefHandle *>(handle);
    if (!h) return 1;
    const size_t nr = h->gl->getNodes().size(), ns = h->sph->getx().size();
    std::memcpy(gl_nodes, h->gl->getNodes().data(), nr * sizeof(double));
    std::memcpy(gl_wts, h->gl->getWeights().data(), nr * sizeof(double));
    std::memcpy(sx, h->sph->getx().data(), ns * sizeof(double));
    std::memcpy(sy, h->sph->gety().data(), ns * sizeof(double));
    std::memcpy(sz, h->sph->getz().data(), ns * sizeof(double));
    std::memcpy(sw, h->sph->getWeights().data(), ns * sizeof(double));
    return 0;
}

void bfsm_ref_destroy(void *handle) { delete static_cast<RefHandle *>(handle); }

} // extern "C"

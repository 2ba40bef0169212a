// B200BoltzmannOperator.hpp -- BoltzmannOperator<B200_Backend>: the drop-in C++ operator class.
//
// Sits next to the reference's two backends
//   BoltzmannOperator<FFTW_Backend>  (Collisions/FFTWBoltzmannOperator.hpp:25-120)
//   BoltzmannOperator<CUDA_Backend>  (Collisions/CUDABoltzmannOperator.hpp:43-131)
// behind the same interface (Collisions/AbstractCollisionOperator.hpp:7-26,
// Collisions/BoltzmannOperator.hpp:7-8): identical constructor signature, initialize(),
// getBackendName(), computeCollision(Q, f_in), operator()(Q, f_in), virtual destructor.
// All device work goes through the C ABI of include/bfsm_b200.h (libbfsm_b200.so).
//
// Pointer convention: like the CUDA backend, Q and f_in are DEVICE pointers and the result is
// complete when computeCollision returns (the reference ends with cudaDeviceSynchronize,
// CUDABoltzmannOperator.cu:218).  Call setHostPointers(true) to get the FFTW backend's
// convention (host pointers; copies happen inside).
//
// Error convention: the reference CUDA backend prints and exits on CUDA errors
// (CUDABoltzmannOperator.hpp:20-38); this class throws std::runtime_error with the message of
// bfsm_last_error() instead, and its destructor is safe on an un-initialised object.
//
// Build inside the reference tree with -DBFSM_HAVE_REFERENCE_HEADERS (uses the reference's own
// base class and quadrature classes); standalone it uses the equivalents in bfsm_host.hpp.
#ifndef B200_BOLTZMANN_OPERATOR_HPP
#define B200_BOLTZMANN_OPERATOR_HPP

#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#ifdef BFSM_HAVE_REFERENCE_HEADERS
#include "Collisions/AbstractCollisionOperator.hpp"
#include "Collisions/BoltzmannOperator.hpp"
#include "Quadratures/GaussLegendre.hpp"
#include "Quadratures/SphericalDesign.hpp"
#else
#include "bfsm_host.hpp"
#endif
#include "bfsm_b200.h"

struct B200_Backend {};

template <> class BoltzmannOperator<B200_Backend> : public AbstractCollisionOperator {
public:
    BoltzmannOperator(std::shared_ptr<GaussLegendreQuadrature> gl_quadrature,
                      std::shared_ptr<SphericalQuadrature> spherical_quadrature, int Nvx, int Nvy,
                      int Nvz, double gamma, double b_gamma, double L)
        : Nvx(Nvx), Nvy(Nvy), Nvz(Nvz), gamma(gamma), b_gamma(b_gamma), L(L),
          gl_quadrature(gl_quadrature), spherical_quadrature(spherical_quadrature)
    {
    }

    // Optional knobs, to be set before initialize().
    void setDevice(int dev) { device = dev; }
    void setHostPointers(bool on) { host_pointers = on; }
    void setStream(void *cuda_stream) { stream = cuda_stream; }
    // Multi-GPU: this operator evaluates pair shard `index` of `count`; give it the
    // rank's communicator (bfsm_comm_init_rank / bfsm_comm_init_all / bfsm_comm_adopt of an ncclComm_t)
    // and computeCollision() ends with the one all-reduce, leaving Q(f,f) on every rank.  Operators of
    // several GPUs driven by ONE host thread must be called through computeCollisionGroup().
    void setShard(int index, int count)
    {
        shard_index = index;
        shard_count = count;
    }
    void setCommunicator(bfsm_comm *c) { comm = c; }
    // Tuning options (bfsm_plan_options); defaults otherwise.
    void setOptions(const bfsm_plan_options &o)
    {
        options = o;
        have_options = true;
    }

    void initialize() override
    {
        if (plan) return;
        const std::vector<double> &rho = gl_quadrature->getNodes();
        const std::vector<double> &w_r = gl_quadrature->getWeights();
        const std::vector<double> &sx = spherical_quadrature->getx();
        const std::vector<double> &sy = spherical_quadrature->gety();
        const std::vector<double> &sz = spherical_quadrature->getz();
        const std::vector<double> &w_s = spherical_quadrature->getWeights();
        check(bfsm_plan_create_ex(&plan, Nvx, Nvy, Nvz, (int)rho.size(), rho.data(), w_r.data(),
                                  (int)sx.size(), sx.data(), sy.data(), sz.data(), w_s.data(), gamma,
                                  b_gamma, L, device, shard_index, shard_count, 0u,
                                  have_options ? &options : nullptr));
    }

    std::string getBackendName() const override { return "B200"; }

    void computeCollision(double *Q, const double *f_in) override
    {
        if (!plan) throw std::runtime_error("BoltzmannOperator<B200_Backend>: initialize() not called");
        if (shard_count > 1) {
            if (!comm) throw std::runtime_error("BoltzmannOperator<B200_Backend>: a sharded operator needs setCommunicator()");
            if (host_pointers) throw std::runtime_error("BoltzmannOperator<B200_Backend>: sharded evaluation takes device pointers");
            check(bfsm_collide_sharded(plan, comm, Q, f_in, stream));
            check(bfsm_sync(plan, stream));
        } else if (host_pointers) {
            check(bfsm_collide_host(plan, Q, f_in, 1, stream));
        } else {
            check(bfsm_collide(plan, Q, f_in, 1, stream));
            check(bfsm_sync(plan, stream)); // result valid on return, like the CUDA backend
        }
    }

    // One host thread, one sharded operator per GPU (communicators from bfsm_comm_init_all): every
    // rank's kernels are enqueued, the all-reduces go out as one NCCL group, all streams are synchronised.
    static void computeCollisionGroup(const std::vector<BoltzmannOperator<B200_Backend> *> &ops,
                                      const std::vector<double *> &Q, const std::vector<const double *> &f_in)
    {
        const int n = (int)ops.size();
        std::vector<bfsm_plan *> plans(n);
        std::vector<bfsm_comm *> comms(n);
        std::vector<double *> q(Q);
        std::vector<const double *> f(f_in);
        std::vector<void *> streams(n);
        for (int k = 0; k < n; ++k) {
            if (!ops[k]->plan || !ops[k]->comm) throw std::runtime_error("computeCollisionGroup: operator not ready");
            plans[k] = ops[k]->plan;
            comms[k] = ops[k]->comm;
            streams[k] = ops[k]->stream;
        }
        check(bfsm_collide_sharded_group(n, plans.data(), comms.data(), q.data(), f.data(), streams.data()));
        for (int k = 0; k < n; ++k) check(bfsm_sync(plans[k], streams[k]));
    }

    // Batch of independent cells (space-inhomogeneous use): n_cells consecutive grids.
    void computeCollisionBatch(double *Q, const double *f_in, int n_cells)
    {
        if (!plan) throw std::runtime_error("BoltzmannOperator<B200_Backend>: initialize() not called");
        if (host_pointers) {
            check(bfsm_collide_host(plan, Q, f_in, n_cells, stream));
        } else {
            check(bfsm_collide(plan, Q, f_in, n_cells, stream));
            check(bfsm_sync(plan, stream));
        }
    }

    void operator()(double *Q, const double *f_in) override { computeCollision(Q, f_in); }

    bfsm_plan *handle() const { return plan; }

    ~BoltzmannOperator() override
    {
        if (plan) bfsm_plan_destroy(plan);
    }

    BoltzmannOperator(const BoltzmannOperator &) = delete;
    BoltzmannOperator &operator=(const BoltzmannOperator &) = delete;

protected:
    const int Nvx, Nvy, Nvz;
    const double gamma, b_gamma, L;
    const std::shared_ptr<GaussLegendreQuadrature> gl_quadrature;
    const std::shared_ptr<SphericalQuadrature> spherical_quadrature;

private:
    static void check(int rc)
    {
        if (rc != BFSM_OK) throw std::runtime_error(std::string("bfsm: ") + bfsm_last_error());
    }
    bfsm_plan *plan = nullptr;
    int device = 0;
    int shard_index = 0, shard_count = 1;
    bfsm_comm *comm = nullptr; // not owned
    bfsm_plan_options options;
    bool have_options = false;
    bool host_pointers = false;
    void *stream = nullptr;
};

#endif // B200_BOLTZMANN_OPERATOR_HPP

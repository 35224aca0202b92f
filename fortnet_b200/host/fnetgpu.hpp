// fnetgpu.hpp -- C++ host-side mirror of the reference's hot-path interface on top of the C ABI
// (include/fnetgpu.h).  The reference is compiled Fortran; with no Fortran compiler in this
// image the host layer a driver would link against is provided in C++ (header-only), with the
// reference's names and argument meaning:
//
//   fnet::TAcsf   ~ type TAcsf  (lib_descriptors/acsf.F90:107-138): calculate / forces (calculatePrime fused)
//   fnet::TBpnn   ~ type TBpnn  (lib_nn/bpnn.F90:46-90): serialWeightsAndBiasesFillup, updateGradients,
//                                predictBatch, loss
//   fnet::TDataset  the hot-path subset of TDataset (lib_io/fnetdata.F90)
//
// Errors follow the reference's abort-on-error convention (lib_dftbp/message.F90:73-103) as
// C++ exceptions carrying fnetgpu_last_error().
#pragma once
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/fnetgpu.h"

namespace fnet {

struct Error : std::runtime_error { using std::runtime_error::runtime_error; };

struct TGFunction {          // acsf.F90:40-75
  int type = FNETGPU_G2;     // FNETGPU_G1..G5
  double rCut = 0, kappa = 0, rs = 0, eta = 0, lambda = 0, xi = 0;
  int atomId = 0;
  int atomicNumbers[2] = {0, 0};
};

struct TDataset {            // flat, concatenated over structures
  std::vector<int> offsets;          // [nStruct+1]
  std::vector<double> coords;        // [3*N] Cartesian Bohr
  std::vector<int> periodic;         // [nStruct]
  std::vector<double> latVecs;       // [9*nStruct], latVecs[9*s+3*k+c] = latVecs(c,k)
  std::vector<int> localAtToAtNum;   // [N]
  std::vector<int> localAtToGlobalSp;// [N], 1-based
  std::vector<int> weights;          // [nStruct]
  std::vector<double> atomicWeights; // [N]
  int nGlobalTargets = 0, nAtomicTargets = 0, nExtFeatures = 0;
  std::vector<double> globalTargets, atomicTargets, extFeatures;
  int nDatapoints() const { return (int)offsets.size() - 1; }
  int nAtoms() const { return offsets.empty() ? 0 : offsets.back(); }
};

class TEnv {                 // one GPU; ~ TEnv_init / destructGlobalEnv
 public:
  explicit TEnv(int device = -1, int precision = 64, bool deterministic = true) {
    if (fnetgpu_init(&ctx_, device, precision, deterministic ? 1 : 0)) throw Error(fnetgpu_last_error(nullptr));
  }
  ~TEnv() { fnetgpu_finalize(ctx_); }
  TEnv(const TEnv &) = delete;
  TEnv &operator=(const TEnv &) = delete;
  fnetgpu_ctx *ctx() const { return ctx_; }
  void check(int rc) const { if (rc) throw Error(fnetgpu_last_error(ctx_)); }
  void upload(int slot, const TDataset &d) {
    check(fnetgpu_dataset_upload(ctx_, slot, d.nDatapoints(), d.offsets.data(), d.coords.data(), d.periodic.data(),
                                 d.latVecs.data(), d.localAtToAtNum.data(), d.localAtToGlobalSp.data(),
                                 d.weights.empty() ? nullptr : d.weights.data(),
                                 d.atomicWeights.empty() ? nullptr : d.atomicWeights.data(), d.nGlobalTargets,
                                 d.globalTargets.data(), d.nAtomicTargets, d.atomicTargets.data(), d.nExtFeatures,
                                 d.extFeatures.data()));
    nAtoms_[slot] = d.nAtoms(); nStruct_[slot] = d.nDatapoints();
  }
  int nAtoms(int slot) const { return nAtoms_[slot]; }
  int nStruct(int slot) const { return nStruct_[slot]; }

 private:
  fnetgpu_ctx *ctx_ = nullptr;
  int nAtoms_[FNETGPU_MAX_SLOTS] = {0}, nStruct_[FNETGPU_MAX_SLOTS] = {0};
};

class TAcsf {
 public:
  // TAcsf_init (acsf.F90:148-167)
  TAcsf(TEnv &env, const std::vector<TGFunction> &functions, bool tZscore) : env_(env), tZscore_(tZscore) {
    const int F = (int)functions.size();
    std::vector<int> type(F), atomid(F), z(2 * F);
    std::vector<double> rc(F), kappa(F), rs(F), eta(F), lam(F), xi(F);
    for (int a = 0; a < F; a++) {
      const TGFunction &g = functions[a];
      type[a] = g.type; rc[a] = g.rCut; kappa[a] = g.kappa; rs[a] = g.rs; eta[a] = g.eta; lam[a] = g.lambda;
      xi[a] = g.xi; atomid[a] = g.atomId; z[2 * a] = g.atomicNumbers[0]; z[2 * a + 1] = g.atomicNumbers[1];
    }
    env_.check(fnetgpu_acsf_set(env_.ctx(), F, type.data(), rc.data(), kappa.data(), rs.data(), eta.data(),
                                lam.data(), xi.data(), atomid.data(), z.data()));
    env_.check(fnetgpu_features_config(env_.ctx(), 0, nullptr));
    nFunc_ = F;
  }
  // TAcsf%calculate (acsf.F90:540-639).  zPrec: [2*F] means then sigmas; empty -> computed and kept.
  void calculate(int slot, const std::vector<double> *zPrecIn = nullptr) {
    if (zPrecIn) zPrec = *zPrecIn;
    const bool have = !zPrec.empty();
    if (!have) zPrec.assign(2 * (size_t)nFunc_, 0.0);
    env_.check(fnetgpu_acsf_calculate(env_.ctx(), slot, tZscore_ ? 1 : 0, zPrec.data(), have ? 1 : 0));
    if (!tZscore_) zPrec.clear();
  }
  std::vector<double> values(int slot) const {   // this%vals%vals(:)%array, concatenated (F, N)
    std::vector<double> out((size_t)nFunc_ * env_.nAtoms(slot));
    env_.check(fnetgpu_features_get(env_.ctx(), slot, out.data()));
    return out;
  }
  std::vector<double> zPrec;
  int nFunctions() const { return nFunc_; }

 private:
  TEnv &env_;
  bool tZscore_;
  int nFunc_ = 0;
};

class TBpnn {
 public:
  // TBpnn_init (bpnn.F90:96-143); activation ids as in fnetgpu.h
  TBpnn(TEnv &env, const std::vector<int> &dims, int nSpecies, int activation) : env_(env), dims_(dims), nSpecies_(nSpecies) {
    env_.check(fnetgpu_net_set(env_.ctx(), nSpecies, (int)dims.size(), dims.data(), activation));
    nTot_ = fnetgpu_ntot(env_.ctx());
  }
  int nTotParams() const { return nTot_; }
  // serialWeightsAndBiasesFillup (bpnn.F90:782-797): weightsAndBiases(nTot, nSpecies)
  void serialWeightsAndBiasesFillup(const std::vector<double> &wb) { env_.check(fnetgpu_params_set(env_.ctx(), wb.data())); }
  // updateGradients + loss (bpnn.F90:277-283, 394-481): returns TDerivs_serialized(resDd) and the loss
  std::vector<double> updateGradients(int slot, int lossId, double &loss) {
    std::vector<double> dd((size_t)nTot_ * nSpecies_);
    env_.check(fnetgpu_grad(env_.ctx(), slot, lossId, nullptr, dd.data(), &loss, nullptr));
    return dd;
  }
  // predictBatch (bpnn.F90:1001-1058): predicts(nOut, N)
  std::vector<double> predictBatch(int slot) {
    std::vector<double> raw((size_t)dims_.back() * env_.nAtoms(slot));
    env_.check(fnetgpu_predict(env_.ctx(), slot, raw.data()));
    return raw;
  }
  // calculatePrime + nJacobian + forceAnalysis_analytical (forces.F90:317-425): forces(3*nOut, N)
  std::vector<double> forces(int slot) {
    std::vector<double> f((size_t)3 * dims_.back() * env_.nAtoms(slot));
    env_.check(fnetgpu_forces(env_.ctx(), slot, f.data()));
    return f;
  }
  // predictForSocketComm (prg_fnet/fortnet.F90:503-609): one MD step of the resident geometry
  void socketStep(int slot, const std::vector<double> &coords, const double *latVecsOrNull, std::vector<double> &globalPrediction,
                  std::vector<double> &atomicPredictions, std::vector<double> &atomicForces) {
    const size_t nOut = dims_.back(), N = env_.nAtoms(slot);
    globalPrediction.resize(nOut * env_.nStruct(slot)); atomicPredictions.resize(nOut * N); atomicForces.resize(3 * nOut * N);
    env_.check(fnetgpu_socket_step(env_.ctx(), slot, coords.data(), latVecsOrNull, globalPrediction.data(),
                                   atomicPredictions.data(), atomicForces.data()));
  }

 private:
  TEnv &env_;
  std::vector<int> dims_;
  int nSpecies_, nTot_ = 0;
};

}  // namespace fnet

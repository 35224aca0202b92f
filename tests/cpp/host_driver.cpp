// host_driver.cpp -- drives the hot path through the C++ host mirror (fortnet_b200/host/fnetgpu.hpp)
// in the order the Fortran driver does (fortnet.F90:111-214): upload -> TAcsf%calculate ->
// serialWeightsAndBiasesFillup -> updateGradients -> predictBatch -> forces.
// Usage: host_driver <dir>   (inputs/outputs are raw little-endian arrays, see tests/test_cpp_host.py)
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "../../fortnet_b200/host/fnetgpu.hpp"

template <typename T>
static std::vector<T> rd(const std::string &p) {
  std::ifstream f(p, std::ios::binary | std::ios::ate);
  if (!f) { std::cerr << "cannot open " << p << "\n"; std::exit(2); }
  size_t n = (size_t)f.tellg() / sizeof(T);
  std::vector<T> v(n);
  f.seekg(0);
  f.read((char *)v.data(), n * sizeof(T));
  return v;
}
template <typename T>
static void wr(const std::string &p, const std::vector<T> &v) {
  std::ofstream f(p, std::ios::binary);
  f.write((const char *)v.data(), v.size() * sizeof(T));
}

int main(int argc, char **argv) {
  if (argc < 2) return 2;
  const std::string d = std::string(argv[1]) + "/";
  try {
    std::vector<int> meta = rd<int>(d + "meta.i32");   // nG nA nExt F nSpecies act lossId zscore L dims...
    const int nG = meta[0], nA = meta[1], nExt = meta[2], F = meta[3], nSpecies = meta[4], act = meta[5],
              lossId = meta[6], zscore = meta[7], L = meta[8];
    std::vector<int> dims(meta.begin() + 9, meta.begin() + 9 + L);
    fnet::TDataset ds;
    ds.offsets = rd<int>(d + "offsets.i32"); ds.coords = rd<double>(d + "coords.f64");
    ds.periodic = rd<int>(d + "periodic.i32"); ds.latVecs = rd<double>(d + "lat.f64");
    ds.localAtToAtNum = rd<int>(d + "atnum.i32"); ds.localAtToGlobalSp = rd<int>(d + "gsp.i32");
    ds.weights = rd<int>(d + "w.i32"); ds.atomicWeights = rd<double>(d + "aw.f64");
    ds.nGlobalTargets = nG; ds.nAtomicTargets = nA; ds.nExtFeatures = nExt;
    ds.globalTargets = rd<double>(d + "gt.f64"); ds.atomicTargets = rd<double>(d + "at.f64");
    ds.extFeatures = rd<double>(d + "ext.f64");
    std::vector<int> fi = rd<int>(d + "fint.i32");       // [F][4]: type atomid z1 z2
    std::vector<double> fp = rd<double>(d + "fpar.f64"); // [F][6]: rcut kappa rs eta lambda xi
    std::vector<fnet::TGFunction> funcs(F);
    for (int a = 0; a < F; a++) {
      funcs[a].type = fi[4 * a]; funcs[a].atomId = fi[4 * a + 1];
      funcs[a].atomicNumbers[0] = fi[4 * a + 2]; funcs[a].atomicNumbers[1] = fi[4 * a + 3];
      funcs[a].rCut = fp[6 * a]; funcs[a].kappa = fp[6 * a + 1]; funcs[a].rs = fp[6 * a + 2];
      funcs[a].eta = fp[6 * a + 3]; funcs[a].lambda = fp[6 * a + 4]; funcs[a].xi = fp[6 * a + 5];
    }
    std::vector<double> wb = rd<double>(d + "wb.f64");

    fnet::TEnv env(-1, 64, true);
    env.upload(0, ds);
    fnet::TAcsf acsf(env, funcs, zscore != 0);
    acsf.calculate(0);
    wr(d + "out_feats.f64", acsf.values(0));
    wr(d + "out_zprec.f64", acsf.zPrec);
    fnet::TBpnn bpnn(env, dims, nSpecies, act);
    bpnn.serialWeightsAndBiasesFillup(wb);
    double loss = 0.0;
    std::vector<double> dd = bpnn.updateGradients(0, lossId, loss);
    wr(d + "out_dd.f64", dd);
    wr(d + "out_loss.f64", std::vector<double>{loss});
    wr(d + "out_raw.f64", bpnn.predictBatch(0));
    wr(d + "out_forces.f64", bpnn.forces(0));
    // error behaviour: unknown activation must be an error, not a fallback
    bool threw = false;
    try { fnet::TBpnn bad(env, dims, nSpecies, 99); } catch (const fnet::Error &) { threw = true; }
    if (!threw) { std::cerr << "unknown activation was accepted\n"; return 3; }
    std::printf("HOST_OK loss=%.12g\n", loss);
  } catch (const fnet::Error &e) {
    std::cerr << "fnetgpu error: " << e.what() << "\n";
    return 1;
  }
  return 0;
}

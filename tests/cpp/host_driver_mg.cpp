// host_driver_mg.cpp -- ONE process drives nDevices GPUs through the multi-GPU layer of the C ABI
// (fnetgpu_mg_*, include/fnetgpu.h): what a non-MPI Fortran driver calls instead of the per-rank
// getStartAndEndIndex / mpifx_allreduce path (lib_common/parallel.F90:23-56, lib_nn/bpnn.F90:455-467).
// Same input files and outputs as host_driver.cpp.   Usage: host_driver_mg <dir> <nDevices>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "../../include/fnetgpu.h"

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
#define CHECK(call)                                                                                   \
  do { if (call) { std::cerr << "fnetgpu error: " << fnetgpu_mg_last_error(mg) << "\n"; return 1; } } while (0)

int main(int argc, char **argv) {
  if (argc < 3) return 2;
  const std::string d = std::string(argv[1]) + "/";
  const int nDev = std::atoi(argv[2]);
  std::vector<int> meta = rd<int>(d + "meta.i32");   // nG nA nExt F nSpecies act lossId zscore L dims...
  const int nG = meta[0], nA = meta[1], nExt = meta[2], F = meta[3], nSpecies = meta[4], act = meta[5], lossId = meta[6],
            zscore = meta[7], L = meta[8];
  std::vector<int> dims(meta.begin() + 9, meta.begin() + 9 + L);
  std::vector<int> offsets = rd<int>(d + "offsets.i32"), periodic = rd<int>(d + "periodic.i32"), atnum = rd<int>(d + "atnum.i32"),
                   gsp = rd<int>(d + "gsp.i32"), w = rd<int>(d + "w.i32");
  std::vector<double> coords = rd<double>(d + "coords.f64"), lat = rd<double>(d + "lat.f64"), aw = rd<double>(d + "aw.f64"),
                      gt = rd<double>(d + "gt.f64"), at = rd<double>(d + "at.f64"), ext = rd<double>(d + "ext.f64");
  std::vector<int> fi = rd<int>(d + "fint.i32");       // [F][4]: type atomid z1 z2
  std::vector<double> fp = rd<double>(d + "fpar.f64"); // [F][6]: rcut kappa rs eta lambda xi
  std::vector<int> type(F), atomid(F), z(2 * F);
  std::vector<double> rc(F), kappa(F), rs(F), eta(F), lam(F), xi(F);
  for (int a = 0; a < F; a++) {
    type[a] = fi[4 * a]; atomid[a] = fi[4 * a + 1]; z[2 * a] = fi[4 * a + 2]; z[2 * a + 1] = fi[4 * a + 3];
    rc[a] = fp[6 * a]; kappa[a] = fp[6 * a + 1]; rs[a] = fp[6 * a + 2]; eta[a] = fp[6 * a + 3]; lam[a] = fp[6 * a + 4]; xi[a] = fp[6 * a + 5];
  }
  std::vector<double> wb = rd<double>(d + "wb.f64");
  const int nStruct = (int)offsets.size() - 1, N = offsets.back();

  fnetgpu_mg *mg = nullptr;
  if (fnetgpu_mg_init(&mg, nDev, 64, 1)) { std::cerr << "fnetgpu error: " << fnetgpu_mg_last_error(nullptr) << "\n"; return 1; }
  if (fnetgpu_mg_device_count(mg) != nDev) return 4;
  CHECK(fnetgpu_mg_dataset_upload(mg, 0, nStruct, offsets.data(), coords.data(), periodic.data(), lat.data(), atnum.data(), gsp.data(),
                                  w.data(), aw.data(), nG, gt.data(), nA, at.data(), nExt, ext.data()));
  CHECK(fnetgpu_mg_acsf_set(mg, F, type.data(), rc.data(), kappa.data(), rs.data(), eta.data(), lam.data(), xi.data(), atomid.data(), z.data()));
  CHECK(fnetgpu_mg_features_config(mg, 0, nullptr));
  std::vector<double> zprec(2 * F, 0.0);
  CHECK(fnetgpu_mg_acsf_calculate(mg, 0, zscore, zprec.data(), 0));
  std::vector<double> feats((size_t)N * F);
  CHECK(fnetgpu_mg_features_get(mg, 0, feats.data()));
  wr(d + "out_feats.f64", feats);
  wr(d + "out_zprec.f64", zprec);
  CHECK(fnetgpu_mg_net_set(mg, nSpecies, L, dims.data(), act));
  CHECK(fnetgpu_mg_params_set(mg, wb.data()));
  const int nTot = fnetgpu_ntot(fnetgpu_mg_context(mg, 0));
  std::vector<double> dd((size_t)nTot * nSpecies), gpred((size_t)nG * nStruct);
  double loss = 0.0, loss2 = 0.0;
  CHECK(fnetgpu_mg_grad(mg, 0, lossId, nullptr, dd.data(), &loss, gpred.data()));
  CHECK(fnetgpu_mg_loss(mg, 0, lossId, &loss2));
  wr(d + "out_dd.f64", dd);
  wr(d + "out_loss.f64", std::vector<double>{loss, loss2});
  wr(d + "out_gpred.f64", gpred);
  std::vector<double> raw((size_t)N * dims[L - 1]), frc((size_t)3 * N * dims[L - 1]);
  CHECK(fnetgpu_mg_predict(mg, 0, raw.data()));
  CHECK(fnetgpu_mg_forces(mg, 0, frc.data()));
  wr(d + "out_raw.f64", raw);
  wr(d + "out_forces.f64", frc);
  int st0, st1, a0, a1, covered = 0;
  for (int k = 0; k < nDev; k++) { if (fnetgpu_mg_shard(mg, 0, k, &st0, &st1, &a0, &a1)) return 5; covered += a1 - a0; }
  if (covered != N) return 6;
  fnetgpu_mg_finalize(mg);
  std::printf("HOST_MG_OK devices=%d loss=%.12g\n", nDev, loss);
  return 0;
}

#!/usr/bin/env python3
"""Converts the reference's own regression fixtures into small .npz files.

Run HERE (the build container), where /root/reference exists; the GPU box has
no reference tree, so tests read only the committed .npz files.

    python tests/golden/make_golden.py

Sources (all under /root/reference/test/prog/fortnet):
  * datasets/*.hdf5                      -> tests/golden/datasets/<name>.npz
  * every case directory whose HSD input has ``ReadNetStats = Yes`` or runs in
    predict/validate mode (i.e. no RANLUX initialisation is involved)
                                         -> tests/golden/cases/<case>.npz
    holding the input netstat (fortnet.hdf5), the golden output netstat
    (_fortnet.hdf5), the golden predictions/forces (_fnetout.hdf5) and the few
    HSD settings that matter to the hot path.

Nothing from the reference is copied except these binary test vectors, which
are the known-answer tests the reference's comparator (bin/testwithworkdir.py,
ATOL 1e-10 / RTOL 1e-9) checks.
"""
import io
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "tools"))
from minih5 import H5File  # noqa: E402

REF = "/root/reference/test/prog/fortnet"


# ---------------------------------------------------------------------------
# minimal HSD parser (nested blocks, "Key = Type { ... }", "Key = value")
# ---------------------------------------------------------------------------
def parse_hsd(text):
    text = re.sub(r"#.*", "", text)
    text = re.sub(r"<<[+<!]*\s*\S+", "", text)  # file includes (socket cases only; skipped)
    toks = re.findall(r"\{|\}|=|'[^']*'|\"[^\"]*\"|[^\s{}=]+", text)
    pos = 0

    def block():
        nonlocal pos
        out = {}
        while pos < len(toks) and toks[pos] != "}":
            key = toks[pos].lower()
            pos += 1
            if toks[pos] == "{":
                pos += 1
                val = block()
                pos += 1
            else:
                assert toks[pos] == "=", (key, toks[pos])
                pos += 1
                vals = []
                # value tokens run until a '{' (typed block) or the next "key =" / "key {" / '}'
                while pos < len(toks) and toks[pos] not in "{}":
                    if pos + 1 < len(toks) and toks[pos + 1] in ("=", "{") and vals:
                        break
                    vals.append(toks[pos].strip("'\""))
                    pos += 1
                    if pos < len(toks) and toks[pos] == "{":
                        break
                if pos < len(toks) and toks[pos] == "{":
                    pos += 1
                    inner = block()
                    pos += 1
                    val = {"_type": " ".join(vals).lower(), **inner}
                else:
                    val = " ".join(vals)
            if key in out:  # repeated keys (e.g. several Function blocks)
                if not isinstance(out[key], list):
                    out[key] = [out[key]]
                out[key].append(val)
            else:
                out[key] = val
        return out

    return block()


def yes(v):
    return str(v).lower() in ("yes", "true", ".true.", "1")


# ---------------------------------------------------------------------------
# dataset / netstat / fnetout converters
# ---------------------------------------------------------------------------
def convert_dataset(path):
    f = H5File(path)
    a = f.attrs("fnetdata/dataset")
    n = int(a["ndatapoints"])
    nExt = int(a.get("nextfeatures", 0))
    withStruct = int(a.get("withstructures", 1))
    tr = f.attrs("fnetdata/dataset/training") if f.exists("fnetdata/dataset/training") else {}
    out = {}
    natoms, coords, lat, periodic, frac = [], [], [], [], []
    atnum, gsp, aw, gt, at, ext, wts = [], [], [], [], [], [], []
    nG = nA = 0
    for i in range(1, n + 1):
        p = "fnetdata/dataset/datapoint%d" % i
        wts.append(int(f.attrs(p).get("weight", 1)))
        if withStruct:
            g = p + "/geometry"
            ga = f.attrs(g)
            c = f[g + "/coordinates"]
            natoms.append(c.shape[0])
            coords.append(c)
            periodic.append(int(ga["periodic"]))
            frac.append(int(ga["fractional"]))
            lat.append(f[g + "/basis"] if periodic[-1] else np.zeros((3, 3)))
            atnum.append(f[g + "/localattoatnum"].astype(np.int32))
            gsp.append(f[g + "/localattoglobalsp"].astype(np.int32))
        if f.exists(p + "/atomicweights"):
            aw.append(f[p + "/atomicweights"])
        else:
            aw.append(np.ones(natoms[-1]))
        if f.exists(p + "/globaltargets"):
            t = f[p + "/globaltargets"]
            nG = t.shape[0]
            gt.append(t)
        if f.exists(p + "/atomictargets"):
            t = f[p + "/atomictargets"]
            nA = t.shape[1]
            at.append(t)
        if nExt > 0 and f.exists(p + "/extfeatures"):
            ext.append(f[p + "/extfeatures"])
    N = int(np.sum(natoms))
    out["natoms"] = np.asarray(natoms, np.int32)
    out["coords"] = np.concatenate(coords).astype(np.float64)            # (N,3) as stored
    out["latvecs"] = np.asarray(lat, np.float64)                           # (nS,3,3) row k = vector k
    out["periodic"] = np.asarray(periodic, np.int32)
    out["fractional"] = np.asarray(frac, np.int32)
    out["atnum"] = np.concatenate(atnum)
    out["globalsp"] = np.concatenate(gsp)
    out["atomicweights"] = np.concatenate(aw).astype(np.float64)
    out["weights"] = np.asarray(wts, np.int32)
    out["globaltargets"] = np.asarray(gt, np.float64).reshape(n, nG) if nG else np.zeros((n, 0))
    out["atomictargets"] = np.concatenate(at).astype(np.float64) if nA else np.zeros((N, 0))
    out["extfeatures"] = np.concatenate(ext).astype(np.float64) if ext else np.zeros((N, 0))
    out["atomicnumbers"] = f["fnetdata/dataset/atomicnumbers"].astype(np.int32)
    assert int(tr.get("nglobaltargets", nG)) == nG and int(tr.get("natomictargets", nA)) == nA
    return out


def convert_netstat(path, prefix):
    """Flattens a netstat file into arrays keyed '<prefix>...'."""
    f = H5File(path)
    out = {}
    meta = {}
    b = "netstat/bpnn"
    ba = f.attrs(b)
    meta["nglobaltargets"] = int(ba["nglobaltargets"])
    meta["natomictargets"] = int(ba["natomictargets"])
    Z = f[b + "/atomicnumbers"].astype(np.int32)
    out[prefix + "atomicnumbers"] = Z
    subnets = {int(f.attrs(b + "/" + k)["element"]): k for k in f.keys(b) if k.endswith("-subnetwork")}
    dims = None
    for isp, z in enumerate(Z):
        sn = b + "/" + subnets[int(z)]
        sa = f.attrs(sn)
        topo = f[sn + "/topology"].astype(np.int32)
        dims = topo
        meta["activation"] = sa["activation"]
        for l in range(1, len(topo)):
            out[prefix + "w_%d_%d" % (isp, l)] = f[sn + "/layer%d/weights" % l]   # (d_{l+1}, d_l)
            out[prefix + "b_%d_%d" % (isp, l)] = f[sn + "/layer%d/bias" % l]      # (d_{l+1},)
    out[prefix + "dims"] = dims
    m = "netstat/mapping"
    if f.exists(m):
        nf = int(f.attrs(m)["nfunctions"])
        funcs = []
        for i in range(1, nf + 1):
            fa = f.attrs(m + "/function%d" % i)
            funcs.append(dict(type=fa["type"], atomid=int(fa["atomid"]), rcut=float(fa["cutoff"]),
                              atomicnumbers=[int(x) for x in fa["atomicnumbers"]],
                              kappa=float(fa.get("kappa", 0.0)), rs=float(fa.get("rs", 0.0)),
                              eta=float(fa.get("eta", 0.0)), lam=float(fa.get("lambda", 0.0)),
                              xi=float(fa.get("xi", 0.0))))
        meta["functions"] = funcs
        if f.exists(m + "/preconditioning"):
            out[prefix + "zmeans"] = f[m + "/preconditioning/means"]
            out[prefix + "zsigmas"] = f[m + "/preconditioning/variances"]
    if f.exists("netstat/external"):
        out[prefix + "extindices"] = f["netstat/external/indices"].astype(np.int32)
    return out, meta


def convert_fnetout(path):
    f = H5File(path)
    o = "fnetout/output"
    a = f.attrs(o)
    n = int(a["ndatapoints"])
    meta = dict(mode=f.attrs("fnetout")["mode"], tforces=int(a.get("tforces", 0)),
                nglobaltargets=int(a["nglobaltargets"]), natomictargets=int(a["natomictargets"]))
    raw, glob, forces = [], [], []
    for i in range(1, n + 1):
        p = o + "/datapoint%d" % i
        if f.exists(p + "/rawpredictions"):
            raw.append(f[p + "/rawpredictions"])
        if f.exists(p + "/globalpredictions"):
            glob.append(f[p + "/globalpredictions"])
        if f.exists(p + "/forces"):
            forces.append(f[p + "/forces"])
    out = {}
    if raw:
        out["out_rawpredictions"] = np.concatenate(raw)      # (N, nOut)
    if glob:
        out["out_globalpredictions"] = np.asarray(glob)       # (nS, nG)
    if forces:
        out["out_forces"] = np.concatenate(forces)            # (N, 3*nG)
    return out, meta


def main():
    os.makedirs(os.path.join(HERE, "datasets"), exist_ok=True)
    os.makedirs(os.path.join(HERE, "cases"), exist_ok=True)
    for name in sorted(os.listdir(os.path.join(REF, "datasets"))):
        if name.endswith(".hdf5"):
            d = convert_dataset(os.path.join(REF, "datasets", name))
            np.savez_compressed(os.path.join(HERE, "datasets", name[:-5] + ".npz"), **d)
    index = []
    with open(os.path.join(REF, "tests")) as fh:
        cases = [l.split()[0] for l in fh if l.strip() and not l.startswith("#")]
    for case in cases:
        cdir = os.path.join(REF, case)
        hsdp = os.path.join(cdir, "fortnet_in.hsd")
        if not os.path.isfile(hsdp):
            continue
        hsd = parse_hsd(open(hsdp).read())
        opt = hsd.get("options", {})
        mode = str(opt.get("mode", "train")).lower()
        readnet = yes(opt.get("readnetstats", "no")) or mode in ("predict", "validate")
        if not readnet or not os.path.isfile(os.path.join(cdir, "fortnet.hdf5")):
            continue
        if "driver" in hsd:
            continue  # socket cases need the i-PI round trip
        data = hsd.get("data", {})
        meta = dict(case=case, mode=mode,
                    dataset=os.path.basename(str(data.get("dataset", "")))[:-5],
                    validset=os.path.basename(str(data.get("validset", "")))[:-5] or None)
        tr = hsd.get("training")
        if isinstance(tr, dict):
            meta["training"] = {k: v for k, v in tr.items() if not isinstance(v, dict)}
            meta["training"]["type"] = tr.get("_type")
            if isinstance(tr.get("regularization"), dict):
                meta["training"]["regularization"] = tr["regularization"]
        an = hsd.get("analysis")
        if isinstance(an, dict) and "forces" in an:
            fo = an["forces"]
            meta["forces"] = fo if isinstance(fo, dict) else {"_type": str(fo).lower()}
        arrays, nmeta = convert_netstat(os.path.join(cdir, "fortnet.hdf5"), "in_")
        meta["netstat"] = nmeta
        if os.path.isfile(os.path.join(cdir, "_fortnet.hdf5")):
            a2, _ = convert_netstat(os.path.join(cdir, "_fortnet.hdf5"), "ref_")
            arrays.update(a2)
        if os.path.isfile(os.path.join(cdir, "_fnetout.hdf5")):
            a3, fmeta = convert_fnetout(os.path.join(cdir, "_fnetout.hdf5"))
            arrays.update(a3)
            meta["fnetout"] = fmeta
        arrays["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
        fname = case.replace("/", "__") + ".npz"
        np.savez_compressed(os.path.join(HERE, "cases", fname), **arrays)
        index.append(dict(case=case, file=fname, mode=mode, dataset=meta["dataset"],
                          training=(meta.get("training") or {}).get("type"),
                          forces=(meta.get("forces") or {}).get("_type")))
    with open(os.path.join(HERE, "index.json"), "w") as fh:
        json.dump(index, fh, indent=1)
    print("datasets:", len(os.listdir(os.path.join(HERE, "datasets"))), "cases:", len(index))
    # ---- fresh-training cases (RANLUX initialisation, no stored input netstat): the golden OUTPUT
    #      netstat + the HSD settings; replayed by tests/test_oracle_fresh_training.py with tests/ranlux.py
    os.makedirs(os.path.join(HERE, "fresh"), exist_ok=True)
    fresh = []
    for case in cases:
        cdir = os.path.join(REF, case)
        hsdp = os.path.join(cdir, "fortnet_in.hsd")
        if not os.path.isfile(hsdp) or not os.path.isfile(os.path.join(cdir, "_fortnet.hdf5")):
            continue
        hsd = parse_hsd(open(hsdp).read())
        opt = hsd.get("options", {})
        mode = str(opt.get("mode", "train")).lower()
        if mode != "train" or yes(opt.get("readnetstats", "no")) or "driver" in hsd:
            continue
        tr = hsd.get("training")
        if not isinstance(tr, dict):
            continue
        data = hsd.get("data", {})
        meta = dict(case=case, mode=mode, fresh=True, seed=int(opt.get("randomseed", 0)),
                    dataset=os.path.basename(str(data.get("dataset", "")))[:-5])
        meta["training"] = {k: v for k, v in tr.items() if not isinstance(v, dict)}
        meta["training"]["type"] = tr.get("_type")
        if isinstance(tr.get("regularization"), dict):
            meta["training"]["regularization"] = tr["regularization"]
        arrays, nmeta = convert_netstat(os.path.join(cdir, "_fortnet.hdf5"), "ref_")
        meta["netstat"] = nmeta
        arrays["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
        fname = case.replace("/", "__") + ".npz"
        np.savez_compressed(os.path.join(HERE, "fresh", fname), **arrays)
        fresh.append(dict(case=case, file=fname, dataset=meta["dataset"], training=meta["training"]["type"],
                          niterations=int(meta["training"].get("niterations", 0)), seed=meta["seed"]))
    with open(os.path.join(HERE, "index_fresh.json"), "w") as fh:
        json.dump(fresh, fh, indent=1)
    print("fresh-training cases:", len(fresh))


if __name__ == "__main__":
    main()

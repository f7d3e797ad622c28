"""Turn the reference's own golden objects into a small fixture the tests can read anywhere.

Run in the build container (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py [/root/reference]

Writes `tests/golden/reference_goldens.npz`.  Only values stored in the reference's `.RData`
files are copied (no code); the arrays are flattened Rle pieces and region tables.  Sources:
  data/genomeData.RData, data/genomeDataRaw.RData, data/genomeInfo.RData,
  data/genomeFstats.RData, data/genomeRegions.RData,
  inst/extdata/chr21/{fstats,regions,coveragePrep,optionsStats}.Rdata
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.rdata import (RObj, as_vector, dframe_columns, is_rle, read_rdata, rle_parts,  # noqa: E402
                          rlist, strings)


def _flat_rle_df(df: RObj, prefix: str, out: dict):
    cols = dframe_columns(df)
    vals, lens, nruns = [], [], []
    for name, col in cols.items():
        v, l = rle_parts(col)
        vals.append(v)
        lens.append(l)
        nruns.append(len(v))
    out[prefix + "_names"] = np.array(list(cols.keys()))
    out[prefix + "_values"] = np.concatenate(vals)
    out[prefix + "_lengths"] = np.concatenate(lens)
    out[prefix + "_nruns"] = np.asarray(nruns, dtype=np.int64)


def _regions(obj: RObj, prefix: str, out: dict):
    parts = rlist(obj)
    gr = parts["regions"]
    rng = gr.attr["ranges"]
    out[prefix + "_start"] = np.asarray(rng.attr["start"].value, dtype=np.int64)
    out[prefix + "_width"] = np.asarray(rng.attr["width"].value, dtype=np.int64)
    out[prefix + "_names"] = np.array(strings(rng.attr["NAMES"]))
    for name, col in dframe_columns(gr.attr["elementMetadata"]).items():
        v = as_vector(col)
        if isinstance(col, RObj) and "levels" in (col.attr if not is_rle(col) else {}):
            lev = strings(col.attr["levels"])
            v = np.array([lev[i - 1] == "TRUE" if i > 0 else False for i in v])
        out[f"{prefix}_{name}"] = np.asarray(v)
    out[prefix + "_nullStats"] = as_vector(parts["nullStats"]).astype(np.float64)
    out[prefix + "_nullWidths"] = as_vector(parts["nullWidths"]).astype(np.int64)
    out[prefix + "_nullPermutation"] = as_vector(parts["nullPermutation"]).astype(np.int64)


def main(ref: str):
    out: dict = {}
    gd = rlist(read_rdata(f"{ref}/data/genomeData.RData")["genomeData"])
    _flat_rle_df(gd["coverage"], "genomeData_cov", out)
    pv, pl = rle_parts(gd["position"])
    out["genomeData_pos_values"], out["genomeData_pos_lengths"] = pv.astype(bool), pl

    raw = rlist(read_rdata(f"{ref}/data/genomeDataRaw.RData")["genomeDataRaw"])
    _flat_rle_df(raw["coverage"], "genomeDataRaw_cov", out)
    assert raw["position"] is None

    info = read_rdata(f"{ref}/data/genomeInfo.RData")["genomeInfo"]
    cols = rlist(info)
    for k in ("gender", "pop"):
        f = cols[k]
        lev = strings(f.attr["levels"])
        out[f"genomeInfo_{k}"] = np.array([lev[i - 1] for i in f.value])
    out["genomeInfo_run"] = np.array(strings(cols["run"]))

    fv, fl = rle_parts(read_rdata(f"{ref}/data/genomeFstats.RData")["genomeFstats"])
    out["genomeFstats_values"], out["genomeFstats_lengths"] = fv, fl

    _regions(read_rdata(f"{ref}/data/genomeRegions.RData")["genomeRegions"], "genomeRegions", out)

    ex = f"{ref}/inst/extdata/chr21"
    fv, fl = rle_parts(read_rdata(f"{ex}/fstats.Rdata")["fstats"])
    out["chr21_fstats_values"], out["chr21_fstats_lengths"] = fv, fl
    _regions(read_rdata(f"{ex}/regions.Rdata")["regions"], "chr21_regions", out)
    prep = rlist(read_rdata(f"{ex}/coveragePrep.Rdata")["prep"])
    assert prep["coverageProcessed"] is None
    out["chr21_prep_mclapplyIndex"] = np.asarray(prep["mclapplyIndex"].value)
    pv, pl = rle_parts(prep["position"])
    out["chr21_prep_pos_values"], out["chr21_prep_pos_lengths"] = pv.astype(bool), pl
    out["chr21_prep_meanCoverage"] = as_vector(prep["meanCoverage"]).astype(np.float64)
    gm = rlist(prep["groupMeans"])
    out["chr21_prep_groupNames"] = np.array(list(gm.keys()))
    for k, v in gm.items():
        out[f"chr21_prep_groupMeans_{k}"] = as_vector(v).astype(np.float64)
    opts = rlist(read_rdata(f"{ex}/optionsStats.Rdata")["optionsStats"])
    models = rlist(opts["models"])
    for k in ("mod", "mod0"):
        m = models[k]
        dim = m.attr["dim"].value
        out[f"chr21_{k}"] = np.asarray(m.value).reshape((dim[1], dim[0])).T.copy()  # column-major
    for k in ("cutoffPre", "scalefac", "cutoffFstat", "nPermute", "seeds", "cutoffFstatUsed"):
        if k in opts and opts[k] is not None:
            out[f"chr21_opt_{k}"] = np.asarray(opts[k].value)
    out["chr21_opt_cutoffType"] = np.array(strings(opts["cutoffType"]))
    ti = read_rdata(f"{ex}/timeinfo.Rdata")["timeinfo"]
    out["chr21_timeinfo"] = np.asarray(ti.value)
    out["chr21_timeinfo_names"] = np.array(strings(ti.attr["names"]))

    dst = os.path.join(HERE, "reference_goldens.npz")
    np.savez_compressed(dst, **out)
    print(f"wrote {dst}: {len(out)} arrays, {os.path.getsize(dst)} bytes")
    for k, v in out.items():
        print(f"  {k:45s} {v.dtype} {v.shape}")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")

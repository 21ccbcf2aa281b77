#!/usr/bin/env python3
"""Generates tests/golden/icrp_import_golden.json: what the REFERENCE's own ICRPPhantomImportPipeline::importPhantom
(R:src/libopendxmc/icrpphantomimportpipeline.cpp:258-351, compiled unmodified into oracle/_ref/opendxmc_ref) returns for
small organ arrays and the organ / media tables of several phantoms, with and without arm removal.

Run in the build container (needs /root/reference and `make ref`).  The table TEXT handed to the reference is the one
workloads.icrp_dat_text() regenerates from the packaged icrp_tables.json (the GPU box has no /root/reference); the script
first checks that the reference parses that text to the same result as its original files."""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
import opendxmc_b200 as dx  # noqa: E402

EXE = os.path.join(ROOT, "oracle", "_ref", "opendxmc_ref")


def reference_import(raw, organs_path, media_path, remove_arms, tmp):
    path = os.path.join(tmp, "organs.bin")
    raw.tofile(path)
    r = subprocess.run([EXE, "icrp", path, organs_path, media_path, str(raw.size), "1", "1", "1" if remove_arms else "0"],
                       capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stderr)
    return json.loads(r.stdout)


def main():
    cases = []
    with tempfile.TemporaryDirectory() as tmp:
        for phantom in ("AM", "AF", "10M", "00F"):
            organs_text, media_text = dx.workloads.icrp_dat_text(phantom)
            op, mp = os.path.join(tmp, "o.dat"), os.path.join(tmp, "m.dat")
            open(op, "w").write(organs_text)
            open(mp, "w").write(media_text)
            orig = "/root/reference/data/phantoms/icrp/%s/%s_" % (phantom, phantom)
            ids = np.array([o["id"] for o in dx.workloads.icrp_tables()[phantom]["organs"]], dtype=np.uint8)
            rng = np.random.default_rng(7)
            arrays = {"all": np.concatenate([ids, ids[::-1], np.zeros(5, dtype=np.uint8)]),
                      "sparse": rng.choice(np.concatenate([ids[::4], [0]]).astype(np.uint8), 300).astype(np.uint8)}
            for nm, raw in arrays.items():
                for remove in (False, True):
                    ref = reference_import(raw, op, mp, remove, tmp)
                    ref_orig = reference_import(raw, orig + "organs.dat", orig + "media.dat", remove, tmp)
                    assert ref == ref_orig, "the regenerated table text does not parse like the original files: " + phantom
                    cases.append({"name": "%s %s arms_removed=%d" % (phantom, nm, remove), "phantom": phantom, "remove_arms": remove,
                                  "input": raw.tolist(), "organ": ref["organ"], "organ_names": ref["organ_names"], "material": ref["material"],
                                  "density": ref["density"], "media_names": [m["name"] for m in ref["materials"]],
                                  "media_composition": [m["Z"] for m in ref["materials"]]})
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "icrp_import_golden.json")
    json.dump({"generator": "tests/golden/make_icrp_golden.py", "source": "oracle/_ref/opendxmc_ref icrp (the reference's importPhantom)",
               "cases": cases}, open(out, "w"), separators=(",", ":"))
    print("wrote", out, len(cases), "cases", os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()

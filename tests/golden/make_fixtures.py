"""Packs the reference's own end-to-end golden fixtures into tests/golden/*.tar.gz.

Run in the build container (needs /root/reference):  python tests/golden/make_fixtures.py
Each archive holds   <case>/sdp/*.json   (the SDP input, JSON form) and   <case>/out/...   (the
reference's outputs: out.txt, iterations*.json, x_*.txt, y.txt, z.txt, c_minus_By/c_minus_By.json)
copied byte for byte from  /root/reference/test/data/end-to-end_tests/<source>/output/ .
Nothing is generated or altered: these are the answers the reference's end-to-end test
(test/src/integration_tests/cases/end-to-end.test.cxx:180-381) compares against at 2^-99.
tests/golden/cases.json records the sdpb command-line options that test uses for each case.
"""
import io
import json
import os
import tarfile

SRC = "/root/reference/test/data/end-to-end_tests"
HERE = os.path.dirname(os.path.abspath(__file__))

COMMON = ("--checkpointInterval 3600 --maxRuntime 1340 --dualityGapThreshold 1.0e-30 "
          "--primalErrorThreshold 1.0e-30 --dualErrorThreshold 1.0e-30 --initialMatrixScalePrimal 1.0e20 "
          "--initialMatrixScaleDual 1.0e20 --feasibleCenteringParameter 0.1 --infeasibleCenteringParameter 0.3 "
          "--stepLengthReduction 0.7 --maxComplementarity 1.0e100 --maxIterations 1000 --verbosity 2 "
          "--procGranularity 1 --writeSolution x,y,z")
ALLOWED = ("--checkpointInterval 3600 --maxRuntime 1341 --dualityGapThreshold 1.0e-30 "
           "--primalErrorThreshold 1.0e-200 --dualErrorThreshold 1.0e-200 --initialMatrixScalePrimal 1.0e20 "
           "--initialMatrixScaleDual 1.0e20 --feasibleCenteringParameter 0.1 --infeasibleCenteringParameter 0.3 "
           "--stepLengthReduction 0.7 --maxComplementarity 1.0e100 --maxIterations 1000 --verbosity 2 "
           "--procGranularity 1 --writeSolution y,z --detectPrimalFeasibleJump --detectDualFeasibleJump "
           "--maxSharedMemory=100.1K")
# name -> (source dir, precision, sdpb args (end-to-end.test.cxx:186-380), iterations file, out.txt keys)
CASES = {
    "1d": ("1d", 664, "", "iterations.json", None),
    "1d-constraints": ("1d-constraints", 768, "", "iterations.json", None),
    "dfibo": ("dfibo-0-0-j=3-c=3.0000-d=3-s=6", 768,
              "--findDualFeasible --findPrimalFeasible --initialMatrixScalePrimal 1e10 "
              "--initialMatrixScaleDual 1e10 --maxComplementarity 1e30 --dualErrorThreshold 1e-10 "
              "--primalErrorThreshold 1e-153 --maxRuntime 259200 --checkpointInterval 3600 --maxIterations 1000 "
              "--feasibleCenteringParameter=0.1 --infeasibleCenteringParameter=0.3 --stepLengthReduction=0.7 "
              "--maxSharedMemory=100K", "iterations.json", None),
    "SingletScalar_cT_test_nmax6": ("SingletScalar_cT_test_nmax6/primal_dual_optimal", 768, COMMON,
                                    "iterations.1.json", None),
    "SingletScalarAllowed_primal_feasible_jump": (
        "SingletScalarAllowed_test_nmax6/primal_feasible_jump", 768, ALLOWED, "iterations.json",
        ["terminateReason", "primalObjective", "dualObjective", "dualityGap", "dualError"]),
    "SingletScalarAllowed_dual_feasible_jump": (
        "SingletScalarAllowed_test_nmax6/dual_feasible_jump", 768, ALLOWED, "iterations.json",
        ["terminateReason", "primalObjective", "dualObjective", "dualityGap", "primalError"]),
}


def main():
    meta = {}
    for name, (src, prec, args, iters, keys) in CASES.items():
        root = os.path.join(SRC, src, "output")
        path = os.path.join(HERE, name + ".tar.gz")
        with tarfile.open(path, "w:gz", compresslevel=9) as tar:
            for sub in ("sdp", "out"):
                for dirpath, _, files in sorted(os.walk(os.path.join(root, sub))):
                    for f in sorted(files):
                        if sub == "sdp" and f == "pmp_info.json":
                            continue  # only spectrum reads it
                        full = os.path.join(dirpath, f)
                        arc = os.path.join(name, os.path.relpath(full, root))
                        data = open(full, "rb").read()
                        info = tarfile.TarInfo(arc)
                        info.size = len(data)
                        info.mtime = 0
                        tar.addfile(info, io.BytesIO(data))
        meta[name] = {"source": "test/data/end-to-end_tests/" + src + "/output", "precision": prec,
                      "sdpb_args": args.split(), "iterations": iters, "out_txt_keys": keys}
        print(name, os.path.getsize(path) // 1024, "KiB")
    json.dump(meta, open(os.path.join(HERE, "cases.json"), "w"), indent=1)
    # the reference's own zipped SDP (written by its pvm2sdp at precision 1024, stored entries with
    # data descriptors as libarchive streams them): byte copy, the fixture of the zip reader
    import shutil
    shutil.copyfile("/root/reference/test/data/sdp.zip", os.path.join(HERE, "sdp.zip"))


if __name__ == "__main__":
    main()

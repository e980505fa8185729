"""Row N4: checkpoints of the solver state (csrc/host/checkpoint.hpp) behind the reference's options
--checkpointDir / --initialCheckpointDir / --checkpointInterval / --noFinalCheckpoint, driven through
the host solver on the CPU oracle.  The reference's own check is `run_sdpb_twice`
(test/src/integration_tests/cases/end-to-end.test.cxx:114-122,313-315, issue #219): the second run
loads the final checkpoint of the first and must reproduce the golden output."""
import ctypes
import filecmp
import json
import os

import pytest

import golden_check
from test_golden_trajectory import CASES, _solve, _unpack


def _args(root, out, ck, case, extra=()):
    return (["--sdpDir", os.path.join(root, "sdp"), "--outDir", out, "--checkpointDir", ck,
             "--precision", str(case["precision"])] + case["sdpb_args"] + list(extra))


def test_second_run_loads_the_final_checkpoint_and_reproduces_the_golden(tmp_path, oracle):
    lib = oracle.load_oracle()
    name = "1d"
    case = CASES[name]
    root = _unpack(name, str(tmp_path))
    out, ck = str(tmp_path / "out"), str(tmp_path / "ck")
    first = _solve(lib, "oracle_solve", _args(root, out, ck, case))
    assert not first["checkpoint_loaded"] and first["checkpoint_generation"] == 1
    meta = json.load(open(os.path.join(ck, "checkpoint.json")))
    assert meta["current"] == 1 and meta["backup"] == 0
    assert os.path.exists(os.path.join(ck, "checkpoint_1_0"))
    second = _solve(lib, "oracle_solve", _args(root, out, ck, case))
    assert second["checkpoint_loaded"] and second["checkpoint_generation"] == 2
    assert second["iterations"] <= 1          # already optimal: terminates before the first step
    assert not os.path.exists(os.path.join(ck, "checkpoint_0_0"))   # the backup generation is removed
    kw = {"iterations_name": None}
    if case["out_txt_keys"]:
        kw["keys"] = tuple(case["out_txt_keys"])
    bad = golden_check.diff_out_dirs(out, os.path.join(root, "out"), **kw)
    assert not bad, bad[:5]


def test_interrupted_run_continues_bit_for_bit(tmp_path, oracle):
    """maxIterations stops a run, the binary checkpoint holds x, X, y, Y exactly, and the restarted
    run ends with the very files the uninterrupted run writes."""
    lib = oracle.load_oracle()
    name = "1d"
    case = CASES[name]
    root = _unpack(name, str(tmp_path))
    ref_out = str(tmp_path / "ref_out")
    whole = _solve(lib, "oracle_solve", _args(root, ref_out, str(tmp_path / "ref_ck"), case, ["--noFinalCheckpoint"]))
    assert not os.path.exists(str(tmp_path / "ref_ck"))
    out, ck = str(tmp_path / "out"), str(tmp_path / "ck")
    cut = [a for a in case["sdpb_args"]]
    part = _solve(lib, "oracle_solve", ["--sdpDir", os.path.join(root, "sdp"), "--outDir", out, "--checkpointDir", ck,
                                        "--precision", str(case["precision"])] + cut + ["--maxIterations", "40"])
    assert part["terminateReason"] == "maxIterations exceeded" and part["iterations"] == 40
    rest = _solve(lib, "oracle_solve", _args(root, out, ck, case))
    assert rest["checkpoint_loaded"]
    assert rest["iterations"] + part["iterations"] == whole["iterations"]
    for f in sorted(os.listdir(ref_out)):
        if f.startswith(("x_", "y.txt", "z.txt")):
            assert filecmp.cmp(os.path.join(ref_out, f), os.path.join(out, f), shallow=False), f


def test_text_checkpoint_from_a_written_solution(tmp_path, oracle):
    """load_text_checkpoint.cxx:6-46: x_*.txt, y.txt, X_matrix_*.txt, Y_matrix_*.txt of an out
    directory are a valid initial checkpoint; the run started there is optimal at once."""
    lib = oracle.load_oracle()
    name = "1d"
    case = CASES[name]
    root = _unpack(name, str(tmp_path))
    out1 = str(tmp_path / "out1")
    args = [a for a in case["sdpb_args"]]
    if "--writeSolution" in args:
        k = args.index("--writeSolution")
        del args[k:k + 2]
    base = ["--sdpDir", os.path.join(root, "sdp"), "--precision", str(case["precision"])] + args
    _solve(lib, "oracle_solve", base + ["--outDir", out1, "--checkpointDir", "", "--writeSolution", "x,y,X,Y"])
    assert os.path.exists(os.path.join(out1, "X_matrix_0.txt"))
    out2 = str(tmp_path / "out2")
    second = _solve(lib, "oracle_solve", base + ["--outDir", out2, "--checkpointDir", "", "--initialCheckpointDir", out1,
                                                 "--writeSolution", "x,y"])
    assert second["checkpoint_loaded"] and second["iterations"] <= 2
    bad = golden_check.diff_out_dirs(out2, os.path.join(root, "out"), iterations_name=None,
                                     **({"keys": tuple(case["out_txt_keys"])} if case["out_txt_keys"] else {}))
    assert not bad, bad[:5]


def test_checkpoint_errors_carry_the_references_texts(tmp_path, oracle):
    lib = oracle.load_oracle()
    name = "1d"
    case = CASES[name]
    root = _unpack(name, str(tmp_path))
    out, ck = str(tmp_path / "out"), str(tmp_path / "ck")
    _solve(lib, "oracle_solve", _args(root, out, ck, case, ["--maxIterations", "3"]))

    def fails(args):
        argv = (ctypes.c_char_p * len(args))(*[a.encode() for a in args])
        buf = ctypes.create_string_buffer(8192)
        assert lib.oracle_solve(len(args), argv, buf, 8192) != 0
        return buf.value.decode()

    # an explicit --initialCheckpointDir must hold a checkpoint (SDPB_Parameters.cxx:186-193)
    msg = fails(_args(root, out, ck, case, ["--initialCheckpointDir", str(tmp_path / "nowhere")]))
    assert "Unable to load checkpoint from directory" in msg
    # another precision: the element images have another length
    other = dict(case, precision=case["precision"] + 128)
    msg = fails(_args(root, out, ck, other))
    assert "binary checkpoint file" in msg
    # truncated file
    path = os.path.join(ck, "checkpoint_1_0")
    data = open(path, "rb").read()
    open(path, "wb").write(data[:len(data) // 2])
    msg = fails(_args(root, out, ck, case))
    assert "Corrupted binary checkpoint file" in msg
    # metadata pointing at a missing generation
    os.remove(path)
    msg = fails(_args(root, out, ck, case))
    assert "Missing checkpoint file" in msg


def _to_binary(lib, src, dst, prec):
    buf = ctypes.create_string_buffer(4096)
    lib.oracle_sdp_to_binary.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t]
    rc = lib.oracle_sdp_to_binary(src.encode(), dst.encode(), prec, buf, 4096)
    assert rc == 0, buf.value.decode()


@pytest.mark.parametrize("name", ["1d", "dfibo"])
def test_binary_block_data_replays_the_golden_trajectory(name, tmp_path, oracle):
    """block_data_<j>.bin (SDP_Block_Data.cxx:32-48, boost_serialization.hxx:17-97; what stock pmp2sdp
    writes by default): the fixture's JSON blocks are rewritten as Boost binary archives
    (csrc/host/block_data_bin.hpp), the JSON is gone from the new directory, and the solver replays
    the reference's golden trajectory from the .bin files -- dfibo has empty odd-parity bases
    (0 x n matrices, whose padding row Elemental serialises too)."""
    lib = oracle.load_oracle()
    case = CASES[name]
    root = _unpack(name, str(tmp_path))
    bin_dir = str(tmp_path / "sdp_bin")
    _to_binary(lib, os.path.join(root, "sdp"), bin_dir, case["precision"])
    names = os.listdir(bin_dir)
    assert any(f.endswith(".bin") for f in names) and not any(f.startswith("block_data") and f.endswith(".json") for f in names)
    head = open(os.path.join(bin_dir, "block_data_0.bin"), "rb").read(48)
    assert head[:8] == (22).to_bytes(8, "little") and head[8:30] == b"serialization::archive"
    assert int.from_bytes(head[32:40], "little") == case["precision"]
    out = str(tmp_path / "out")
    summary = _solve(lib, "oracle_solve", ["--sdpDir", bin_dir, "--outDir", out, "--checkpointDir", "",
                                           "--precision", str(case["precision"])] + case["sdpb_args"])
    kw = {"iterations_name": case["iterations"]}
    if case["out_txt_keys"]:
        kw["keys"] = tuple(case["out_txt_keys"])
    bad = golden_check.diff_out_dirs(out, os.path.join(root, "out"), **kw)
    assert not bad, bad[:5]
    assert summary["iterations"] == len(json.load(open(os.path.join(root, "out", case["iterations"]))))


def test_binary_block_data_rejects_another_precision_and_truncation(tmp_path, oracle):
    lib = oracle.load_oracle()
    case = CASES["1d"]
    root = _unpack("1d", str(tmp_path))
    bin_dir = str(tmp_path / "sdp_bin")
    _to_binary(lib, os.path.join(root, "sdp"), bin_dir, case["precision"])

    def fails(args):
        argv = (ctypes.c_char_p * len(args))(*[a.encode() for a in args])
        buf = ctypes.create_string_buffer(8192)
        assert lib.oracle_solve(len(args), argv, buf, 8192) != 0
        return buf.value.decode()

    base = ["--sdpDir", bin_dir, "--outDir", str(tmp_path / "out"), "--checkpointDir", ""]
    msg = fails(base + ["--precision", str(case["precision"] + 64)])
    assert "Read GMP precision: %d, expected: %d" % (case["precision"], case["precision"] + 64) in msg  # SDP_Block_Data.cxx:42-44
    path = os.path.join(bin_dir, "block_data_0.bin")
    data = open(path, "rb").read()
    open(path, "wb").write(data[:len(data) - 100])
    assert "Unexpected end of binary block data" in fails(base + ["--precision", str(case["precision"])])
    open(path, "wb").write(b"\x00" * 64)
    assert "Not a Boost binary archive" in fails(base + ["--precision", str(case["precision"])])


def test_zipped_sdp_of_the_reference_reads_like_its_unpacked_directory(tmp_path, oracle):
    """`pmp2sdp --zip` stores, never deflates (src/pmp2sdp/Archive_Writer.cxx:10-14); sdpb reads the
    archive in place.  tests/golden/sdp.zip is the reference's own test/data/sdp.zip (libarchive:
    local headers with data descriptors).  Solving from the archive and from the directory Python's
    zipfile unpacks it into must write identical files."""
    import zipfile
    from test_golden_trajectory import GOLDEN
    lib = oracle.load_oracle()
    tmp_root = os.environ.get("TMPDIR", "/tmp")
    before = set(d for d in os.listdir(tmp_root) if d.startswith("sdpb_b200_sdp_"))
    archive = os.path.join(GOLDEN, "sdp.zip")
    unpacked = str(tmp_path / "sdp")
    with zipfile.ZipFile(archive) as z:
        assert all(i.compress_type == zipfile.ZIP_STORED for i in z.infolist())
        z.extractall(unpacked)
    base = ["--precision", "1024", "--checkpointDir", "", "--maxIterations", "60", "--writeSolution", "x,y",
            "--dualityGapThreshold", "1e-30", "--primalErrorThreshold", "1e-30", "--dualErrorThreshold", "1e-30"]
    a = _solve(lib, "oracle_solve", ["--sdpDir", archive, "--outDir", str(tmp_path / "out_zip")] + base)
    b = _solve(lib, "oracle_solve", ["--sdpDir", unpacked, "--outDir", str(tmp_path / "out_dir")] + base)
    assert a["iterations"] == b["iterations"] and a["terminateReason"] == b["terminateReason"]
    for f in ("x_0.txt", "y.txt"):
        assert filecmp.cmp(str(tmp_path / "out_zip" / f), str(tmp_path / "out_dir" / f), shallow=False), f
    # a deflated archive is refused with a pointer to the reference's writer
    deflated = str(tmp_path / "deflated.zip")
    with zipfile.ZipFile(deflated, "w", zipfile.ZIP_DEFLATED) as z:
        for f in os.listdir(unpacked):
            z.write(os.path.join(unpacked, f), f)
    argv = ["--sdpDir", deflated, "--outDir", str(tmp_path / "o")] + base
    cargv = (ctypes.c_char_p * len(argv))(*[x.encode() for x in argv])
    buf = ctypes.create_string_buffer(8192)
    assert lib.oracle_solve(len(argv), cargv, buf, 8192) != 0 and "is compressed" in buf.value.decode()
    # the unpacked copies are gone again, also after the failure
    assert set(d for d in os.listdir(tmp_root) if d.startswith("sdpb_b200_sdp_")) == before

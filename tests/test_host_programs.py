"""The reference-shaped host programs (flash-viterbi_b200/host/*.c) and their run.py driver."""
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, load_golden

HOST = ROOT / "flash-viterbi_b200" / "host"
sys.path.insert(0, str(HOST))

P64 = {"K_STATE": 64, "T_STATE": 50, "obserRouteLEN": 256, "prob": 0.253, "MAX_THREADS": 8, "BeamSearchWidth": 8}


def _write_k64_text(tmp_path):
    import gen_hmm

    A, B, Pi = gen_hmm.make_hmm(64, 50, 0.253, 1)  # same numbers as data_script.py -s 1
    g = load_golden("hmm_k64")
    assert np.array_equal(gen_hmm.as_reference_floats(A), g["A"])
    gen_hmm.write_text(tmp_path / "data", 64, 256, 0.253, A, B, Pi, g["obs"][0])
    return g


def test_run_py_substitution_hits_every_knob(fv):
    """src/run.py's regexes (R:29-47) must all match our host sources, and the result must compile."""
    import run as host_run

    for program in host_run.PROGRAMS:
        src = (HOST / f"{program}.c").read_text()
        p = {"K_STATE": 3965, "T_STATE": 50, "obserRouteLEN": 256, "prob": 0.112, "MAX_THREADS": 1, "BeamSearchWidth": 32}
        out = host_run.substitute(src, p, "./somewhere/", program)
        assert "#define K_STATE 3965" in out and "#define MAX_THREADS 1" in out and 'data_path[] = "./somewhere/"' in out
        assert "prob%.3f" in out and "const float prob = 0.112;" in out
        if "BS" in program:
            assert "const int BeamSearchWidth = 32;" in out
    # the three report lines keep the reference's printf shapes (F:378, F:119-123)
    src = (HOST / "FLASH_Viterbi_multithread.c").read_text()
    for shape in ('"time: %lf \\n"', '"path: ["', '"%d "', '"memory: %d\\n"'):
        assert shape in src


def test_host_programs_compile(fv, tmp_path):
    import run as host_run

    for program in host_run.PROGRAMS:
        binary = host_run.compile_program(program, P64, "./data/", tmp_path)
        assert binary.exists()


@pytest.mark.gpu
def test_host_program_matches_reference_binary(fv, tmp_path):
    """Our program and the UNMODIFIED reference binary on the same text files: byte-identical
    path: and memory: lines (the reference binaries are prebuilt into oracle/_ref/)."""
    import run as host_run
    from oracle import build_ref

    g = _write_k64_text(tmp_path)
    for program, prog, B in (("FLASH_Viterbi_multithread", "FLASH", None), ("FLASH_BS_Viterbi_multithread", "FLASH_BS", 8)):
        binary = host_run.compile_program(program, P64, "./data/", tmp_path / "build")
        ours = subprocess.run([str(binary)], cwd=tmp_path, capture_output=True, text=True)
        assert ours.returncode == 0, ours.stderr
        assert re.search(r"time: ([\d.]+)", ours.stdout) and re.search(r"memory: (\d+)", ours.stdout)  # R:75-76
        ref_bin = build_ref.OUT_DIR / build_ref.binary_name(prog, 64, 50, 256, 0.253, 8, B)
        if not ref_bin.exists():
            if not build_ref.reference_available():
                pytest.skip("reference binary not prebuilt")
            ref_bin = build_ref.build(prog, 64, 50, 256, 0.253, 8, B)
        ref = subprocess.run([str(ref_bin)], cwd=tmp_path, capture_output=True, text=True)
        assert ref.returncode == 0
        pick = lambda out, key: [ln for ln in out.splitlines() if ln.startswith(key)][0]
        assert pick(ours.stdout, "path:") == pick(ref.stdout, "path:")
        assert pick(ours.stdout, "memory:") == pick(ref.stdout, "memory:")
        # and the golden vector recorded at build time says the same
        row = [c for c in range(len(g["case_prog"])) if g["case_prog"][c] == (0 if B is None else 1)
               and g["case_seq"][c] == 0 and g["case_N"][c] == 8 and g["case_B"][c] == (B or 0)][0]
        assert pick(ours.stdout, "path:") == "path: [" + "".join(f"{x} " for x in g["case_path"][row]) + "]"


@pytest.mark.gpu
def test_host_programs_read_dag_named_files(fv, tmp_path):
    """files named as generate_data/data_script_dag.py names them (`*_K{K}_T{T}_DAG.txt`, DD:63-66).
    The reference programs cannot open those names (F:51); ours fall back to them.  Same text the golden
    run decoded (tests/golden/make_golden_dag.py), so the path must equal what the UNMODIFIED reference
    printed for the `prob`-named copies — and, where its binary travelled along, what it prints now."""
    import shutil

    import run as host_run
    from oracle import build_ref

    g = load_golden("dag_k96")
    K, M, T = 96, 50, 64
    data = tmp_path / "data"
    data.mkdir()
    np.savetxt(data / f"A_K{K}_T{T}_DAG.txt", g["A64"], fmt="%.16f")
    np.savetxt(data / f"B_K{K}_T{T}_DAG.txt", g["B64"], fmt="%.16f")
    np.savetxt(data / f"Pi_K{K}_T{T}_DAG.txt", g["Pi64"], fmt="%.16f", newline=" ")
    np.savetxt(data / f"ob_K{K}_T{T}_DAG.txt", g["obs"][0], fmt="%d", newline=" ")
    pick = lambda out, key: [ln for ln in out.splitlines() if ln.startswith(key)][0]
    for program, prog, N, B in (("FLASH_Viterbi_multithread", "FLASH", 9, None), ("FLASH_BS_Viterbi_multithread", "FLASH_BS", 4, 16)):
        p = {"K_STATE": K, "T_STATE": M, "obserRouteLEN": T, "prob": 0.9, "MAX_THREADS": N, "BeamSearchWidth": B or 8}
        binary = host_run.compile_program(program, p, "./data/", tmp_path / "build")
        ours = subprocess.run([str(binary)], cwd=tmp_path, capture_output=True, text=True)
        assert ours.returncode == 0, ours.stderr
        row = [c for c in range(len(g["case_prog"])) if g["case_prog"][c] == (0 if B is None else 1)
               and g["case_N"][c] == N and g["case_B"][c] == (B or 0)][0]
        assert pick(ours.stdout, "path:") == "path: [" + "".join(f"{x} " for x in g["case_path"][row]) + "]"
        assert pick(ours.stdout, "memory:") == f"memory: {g['case_memory'][row]}"
        ref_bin = build_ref.OUT_DIR / build_ref.binary_name(prog, K, M, T, 0.9, N, B)
        if ref_bin.exists():
            for kind in ("A", "B", "Pi", "ob"):
                shutil.copy(data / f"{kind}_K{K}_T{T}_DAG.txt", build_ref.data_file(data, kind, K, T, 0.9))
            ref = subprocess.run([str(ref_bin)], cwd=tmp_path, capture_output=True, text=True)
            assert ref.returncode == 0
            assert pick(ours.stdout, "path:") == pick(ref.stdout, "path:")
            for kind in ("A", "B", "Pi", "ob"):  # the next program must find the _DAG names only
                build_ref.data_file(data, kind, K, T, 0.9).unlink()


@pytest.mark.gpu
def test_run_py_writes_the_reference_csv(fv, tmp_path, monkeypatch):
    """host/run.py end to end, as src/run.py would be used (R:95-107): substitute, compile, run both programs on the
    K=64 data set, append one row per run to result/<program>_result.csv with the reference's nine columns first
    (R:105) and the extra ones after; a second invocation appends without repeating the header."""
    import csv

    import run as host_run

    g = _write_k64_text(tmp_path)
    monkeypatch.setattr(host_run, "parameters", [dict(P64), dict(P64, MAX_THREADS=3, BeamSearchWidth=16)])
    monkeypatch.chdir(tmp_path)
    argv = ["run.py", "--data", "./data/", "--result", str(tmp_path / "result"), "--build-dir", str(tmp_path / "build_host")]
    for _ in range(2):
        monkeypatch.setattr(sys, "argv", argv)
        host_run.main()
    for program in host_run.PROGRAMS:
        rows = list(csv.reader(open(tmp_path / "result" / f"{program}_result.csv", encoding="utf-8")))
        assert rows[0] == host_run.HEADER + host_run.EXTRA and rows[0][:9] == ["timestamp", "K_STATE", "T_STATE", "obserRouteLEN",
                                                                              "prob", "MAX_THREADS", "BeamSearchWidth", "time", "memory"]
        assert len(rows) == 1 + 4  # two parameter sets, two invocations, one header
        for row in rows[1:]:
            d = dict(zip(rows[0], row))
            assert d["K_STATE"] == "64" and d["obserRouteLEN"] == "256" and float(d["time"]) >= 0 and int(d["memory"]) > 0
            assert float(d["time_including_prep"]) >= float(d["time"]) and int(d["executed_steps"]) > 0
        # the memory column is the reference's formula: the golden run recorded the same number
        row8 = [c for c in range(len(g["case_prog"])) if g["case_prog"][c] == (1 if "BS" in program else 0)
                and g["case_seq"][c] == 0 and g["case_N"][c] == 8 and g["case_B"][c] == (8 if "BS" in program else 0)][0]
        assert int(rows[1][8]) == int(g["case_memory"][row8])

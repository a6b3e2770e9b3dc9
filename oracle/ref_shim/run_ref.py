"""run_ref.py — builds (if needed) and runs oracle/_ref/ref_driver, i.e. the reference's own ExodusIO.hpp, on one
mesh and returns what it computed.  TEST INFRASTRUCTURE (oracle/ref_shim/README.md).

    from run_ref import run_reference
    out = run_reference("/root/reference/data/bolted_bracket.exo", nparts=2)
    out["assemble"]["A_vals"], out["solution"]["vals_nod_var1_step2"], out["getmatrix"]["A_cols"], ...

Needs /root/reference (the header is compiled where it lies), so it only works in the build container; the
tests that travel to the GPU box use the fixtures tests/golden/make_ref_golden.py makes from these outputs.
"""
import os
import subprocess
import tempfile

import dump_exo

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE = os.path.dirname(HERE)
DRIVER = os.path.join(ORACLE, "_ref", "ref_driver")
REFERENCE = "/root/reference"


def available() -> bool:
    return os.path.exists(os.path.join(REFERENCE, "ExodusIO.hpp"))


def build(force: bool = False) -> str:
    """(re)builds oracle/_ref/ref_driver where the reference's sources exist; elsewhere (the GPU box) the prebuilt
    binary that travelled with the repo is used as it is"""
    if available() and (force or not os.path.exists(DRIVER)):
        subprocess.run(["make", "-C", ORACLE, "ref"] + (["-B"] if force else []), check=True, capture_output=True)
    if not os.path.exists(DRIVER):
        raise FileNotFoundError(f"{DRIVER} is not built and /root/reference is absent")
    return DRIVER


def run_reference(exo_path: str, nparts: int = 2, get_matrix: bool = True, timeout: int = 600, arrays=None) -> dict:
    """arrays = (x, y, z, blocks, nodesets) (see dump_exo.dump_arrays) runs the reference on a mesh held in memory;
    exo_path is then only the name the reference is given."""
    build()
    with tempfile.TemporaryDirectory() as tmp:
        if arrays is None:
            dump_exo.dump(exo_path, tmp)
        else:
            dump_exo.dump_arrays(os.path.basename(exo_path), tmp, *arrays)
        prefix = os.path.join(tmp, "out")
        cmd = [DRIVER, exo_path, prefix, str(nparts)] + ([] if get_matrix else ["--no-getmatrix"])
        p = subprocess.run(cmd, env=dict(os.environ, REF_SHIM_DUMP_DIR=tmp), capture_output=True, text=True, timeout=timeout)
        out = {"returncode": p.returncode, "stderr": p.stderr[-2000:]}
        for key, suffix in (("assemble", ".assemble.dump"), ("solution", ".solution.exo.shimdump"), ("getmatrix", ".getmatrix.dump")):
            if os.path.exists(prefix + suffix):
                out[key] = dump_exo.load(prefix + suffix)
        if os.path.exists(prefix + ".timing"):
            with open(prefix + ".timing") as f:
                out["timing"] = {k: float(v) for k, v in (line.split() for line in f if line.strip())}
        if os.path.exists(prefix + ".mpi-proc-0.out"):
            with open(prefix + ".mpi-proc-0.out") as f:
                out["dump_text"] = f.read()
        return out

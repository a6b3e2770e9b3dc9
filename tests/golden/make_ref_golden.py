"""Regenerates tests/golden/ref_pins.json by RUNNING THE REFERENCE'S OWN CODE (run in the build container, where
/root/reference exists): oracle/_ref/ref_driver = /root/reference/ExodusIO.hpp compiled unmodified against the
single-rank stand-ins of oracle/ref_shim/ (see its README.md), executed in the call order of the reference's
main() on every mesh under the reference's data/ directory.

Pinned per mesh: what IO::assemble returned (A, B, the reduced->original id map, the nodeset cache), the text its
printCrsMatrix / printMultiVector write for them, what
IO::getMatrix returned on one rank, and everything IO::decompose + IO::writeSolution handed to the Exodus API
for 2 and 4 partitions — as sha256 digests of the arrays (full arrays for the tiny meshes).

    python tests/golden/make_ref_golden.py
"""
import glob
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, "..", "..")
sys.path.insert(0, os.path.join(ROOT, "oracle", "ref_shim"))
import pins as P      # noqa: E402
import run_ref as R   # noqa: E402

DATA = "/root/reference/data"
PARTS = (2, 4)


def pins_for(path: str) -> dict:
    entry = {"decompose": {}}
    for nparts in PARTS:
        out = R.run_reference(path, nparts, get_matrix=(nparts == PARTS[0]))
        if "assemble" not in out:
            raise RuntimeError(f"{path}: the reference's assemble did not finish: {out['stderr']}")
        if nparts == PARTS[0]:
            entry["assemble"] = P.summ_assemble(out["assemble"])
            entry["dump"] = P.summ_dump_text(out["dump_text"])
            entry["getmatrix"] = P.summ_getmatrix(out["getmatrix"]) if "getmatrix" in out else {"failed": out["stderr"].strip().splitlines()[:1]}
        if out["returncode"] == 0 and "solution" in out:
            entry["decompose"][str(nparts)] = P.summ_output(P.canon_from_shimdump(out["solution"]))
        else:   # e.g. initialguess.exo: element type "TET4" is rejected by the reference's decompose
            entry["decompose"][str(nparts)] = {"failed": [l for l in out["stderr"].strip().splitlines() if "ref_driver" not in l][:1]}
    return entry


CUBES = [(5, 5, 5), (6, 4, 3), (3, 3, 9), (9, 9, 9), (17, 17, 17), (33, 17, 9)]
# tiny meshes that exercise the corners of IO::assemble (one TETRA block, coordinates irrelevant): name -> (num_nodes, conn, nodesets)
EDGE = {
    "d3_single_dof": (5, [[0, 1, 2, 3], [0, 2, 3, 4]], {7: [0, 2, 3, 4]}),                  # the DOF's neighbours are all Dirichlet (D3)
    "d3_two_dofs": (6, [[0, 1, 2, 3], [2, 3, 4, 5]], {7: [0, 2, 3, 5]}),
    "isolated_node": (6, [[0, 1, 3, 4], [1, 3, 4, 5]], {9: [5]}),                           # node 2 is in no element
    "empty_nodeset": (5, [[0, 1, 2, 3], [1, 2, 3, 4]], {3: [], 9: [4]}),
    "overlapping_nodesets": (6, [[0, 1, 2, 3], [1, 2, 3, 4], [2, 3, 4, 5]], {5: [0, 5], 2: [5], 8: [0]}),   # D2
}


def synthetic_pins() -> dict:
    """the reference on meshes held in memory: the Kuhn tet cubes of the benchmark family (from
    the oracle's explicit cube mesh) and the edge cases above"""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    O.build()
    syn = {}
    for dims in CUBES:
        m = O.cube_mesh(*dims)
        out = R.run_reference("cube.exo", 2, get_matrix=False, arrays=(m.x, m.y, m.z, [("TETRA", m.conn)], m.nodesets))
        assert out["returncode"] == 0, out["stderr"]
        syn["cube:%dx%dx%d" % dims] = {"assemble": P.summ_assemble(out["assemble"]),
                                       "decompose": {"2": P.summ_output(P.canon_from_shimdump(out["solution"]))}}
    for name, (N, conn, ns) in EDGE.items():
        x = np.arange(N, dtype=np.float64)
        out = R.run_reference(name + ".exo", 2, get_matrix=False, arrays=(x, 0 * x, 0 * x, [("TETRA", np.asarray(conn, dtype=np.int32))], ns))
        assert out["returncode"] == 0, out["stderr"]
        syn["edge:" + name] = {"mesh": {"num_nodes": N, "conn": conn, "nodesets": {str(k): v for k, v in ns.items()}},
                               "assemble": P.summ_assemble(out["assemble"]),
                               "decompose": {"2": P.summ_output(P.canon_from_shimdump(out["solution"]))}}
    return syn


def main():
    R.build(force=True)
    files = sorted(f for f in glob.glob(os.path.join(DATA, "*.exo")) if not f.endswith(".ref.exo"))
    gold = {"_how": "tests/golden/make_ref_golden.py: /root/reference/ExodusIO.hpp run on one rank through oracle/_ref/ref_driver",
            "_solution_stand_in": "x[row] = 0.25 + 0.5*row + 4096*step for writeSolution(X, step), step = 0, 1"}
    for f in files:
        name = os.path.basename(f)[:-4]
        gold[name] = pins_for(f)
        a = gold[name]["assemble"]["A"]
        print(f"{name}: n={a['n']} rows={a['nrows']} nnz={a['nnz']} trace={a['trace']:.0f} sumB={gold[name]['assemble']['sum_B']:.0f}")
    gold["_synthetic"] = synthetic_pins()
    print("synthetic:", ", ".join(gold["_synthetic"]))
    with open(os.path.join(HERE, "ref_pins.json"), "w") as fp:
        # one mesh per line: compact, and a regenerated fixture diffs mesh by mesh
        fp.write("{\n" + ",\n".join(f"{json.dumps(k)}: {json.dumps(gold[k], sort_keys=True)}" for k in sorted(gold)) + "\n}\n")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Generates csrc/kuhn_rows_generated.cuh: the P1 row of a Kuhn-cube node with the element arithmetic PARTIALLY
EVALUATED.  The 24 tets around a node have axis-aligned edges, so most operands of tet_G / K_ab (assemble.cu; the
oracle's tet_G in oracle/heat_oracle.c) are exact zeros.  This script runs the very same operation sequence on symbols,
applying only identities that are exact in IEEE-754 round-to-nearest arithmetic for finite operands:

    x - x = +0        x * 0 = (+-)0        x + (+-)0 = x        0 - x = -x        -(a*b) = (-a)*b       |-x| = |x|
    a*b = b*a         a+b = b+a            -(a+b) = (-a)+(-b)

so every NON-ZERO value it emits is bit-identical to what the unspecialised code computes; zero contributions are
dropped (adding +-0 to an accumulator that started at +0 leaves it unchanged, and x + (-x) = +0 either way).  Common
subexpressions are shared across the 24 tets of a row.  tests/test_gpu_parity.py::test_cube_direct_sell_assembly holds
the result to the oracle bit for bit (zero signs included).

    python tools/gen_kuhn_rows.py > domain-decomposed-pde-solver_b200/csrc/kuhn_rows_generated.cuh
"""
import sys

PERMS = [(0, 1, 2), (0, 2, 1), (1, 0, 2), (1, 2, 0), (2, 0, 1), (2, 1, 0)]

# ---- symbolic values: None = exact zero; otherwise (sign, node) with node a hash-consed tuple ----------------------
nodes = {}      # structural key -> id
defs = []       # id -> key


def mk(key):
    if key not in nodes:
        nodes[key] = len(defs)
        defs.append(key)
    return nodes[key]


def sym(name):
    return (1, mk(("sym", name)))


def neg(a):
    return None if a is None else (-a[0], a[1])


def mul(a, b):
    if a is None or b is None:
        return None
    lo, hi = sorted((a[1], b[1]))
    return (a[0] * b[0], mk(("mul", lo, hi)))


def add(a, b):
    if a is None:
        return b
    if b is None:
        return a
    if a[1] == b[1] and a[0] == -b[0]:
        return None                                  # x + (-x) = +0
    # canonical: factor the sign out of the first (lower id) operand: s*(x + t*y)
    (sa, ia), (sb, ib) = sorted((a, b), key=lambda v: v[1])
    return (sa, mk(("add", ia, sa * sb, ib)))      # value = sa * (node_ia + (sa*sb) * node_ib)


def sub(a, b):
    return add(a, neg(b))


def fabs(a):
    return None if a is None else (1, a[1])


def rcp6(det):                                       # 1.0 / (6.0 * fabs(det))
    return (1, mk(("rcp6", fabs(det)[1])))


def scale(a, c):                                     # a * constant c (1000.0 / 100.0)
    return None if a is None else (a[0], mk(("mulc", a[1], c)))


def tet(p):
    """oracle's tet_G: returns G[4][3], s"""
    a = [sub(p[1][d], p[0][d]) for d in range(3)]
    b = [sub(p[2][d], p[0][d]) for d in range(3)]
    c = [sub(p[3][d], p[0][d]) for d in range(3)]
    cross = lambda u, w: [sub(mul(u[1], w[2]), mul(u[2], w[1])), sub(mul(u[2], w[0]), mul(u[0], w[2])), sub(mul(u[0], w[1]), mul(u[1], w[0]))]
    G = [None, cross(b, c), cross(c, a), cross(a, b)]
    G[0] = [neg(add(add(G[1][d], G[2][d]), G[3][d])) for d in range(3)]
    det = add(add(mul(a[0], G[1][0]), mul(a[1], G[1][1])), mul(a[2], G[1][2]))
    return G, rcp6(det)


def build():
    """-> blocks: [(guard text, [(slot, di, kab)])] in ascending element id; expression nodes live in `defs`"""
    nodes.clear(); defs.clear()                      # node numbering (hence the emitted names) starts afresh
    X = [sym(f"X[{t}]") for t in range(3)]
    Y = [sym(f"Y[{t}]") for t in range(3)]
    Z = [sym(f"Z[{t}]") for t in range(3)]
    coords = (X, Y, Z)
    blocks = []                                      # (guard, [(target, expr, kind)]) in ascending element id
    for oz, kguard in ((1, "klo"), (0, "khi")):
        for oy, jguard in ((1, "jlo"), (0, "jhi")):
            stmts = []
            for ox in (1, 0):
                o = (ox, oy, oz)
                for pm, perm in enumerate(PERMS):
                    V = [[0, 0, 0]]
                    for q in range(3):
                        nxt = list(V[-1]); nxt[perm[q]] += 1; V.append(nxt)
                    if list(o) not in V:
                        continue
                    ai = V.index(list(o))
                    p = [[coords[d][1 - o[d] + V[q][d]] for d in range(3)] for q in range(4)]
                    G, s = tet(p)
                    for q in range(4):
                        kab = mul(add(add(mul(G[ai][0], G[q][0]), mul(G[ai][1], G[q][1])), mul(G[ai][2], G[q][2])), s)
                        if kab is None:
                            continue
                        di, dj, dk = (V[q][d] - o[d] for d in range(3))
                        code = (di != 0) + 2 * (dj != 0) + 4 * (dk != 0)
                        slot = 7 + code if di + dj + dk > 0 else 7 - code
                        stmts.append((slot, di, kab))
            blocks.append((f"{kguard} && {jguard}", stmts))
    return blocks


def evaluate(X, Y, Z, jlo, jhi, klo, khi, at_lo, at_hi):
    """The generated row in Python floats (IEEE doubles, the same operations in the same order as the emitted CUDA):
    -> (v[15], bsum).  tests/test_host_cpu.py holds it to the oracle's rows bit for bit, so the symbolic transformation is
    checked on the CPU, independently of any GPU run."""
    blocks = build()
    env = {"klo": klo, "khi": khi, "jlo": jlo, "jhi": jhi}
    vals = {}

    def val(i):
        if i in vals:
            return vals[i]
        k = defs[i]
        if k[0] == "sym":
            r = {"X": X, "Y": Y, "Z": Z}[k[1][0]][int(k[1][2])]
        elif k[0] == "mul":
            r = val(k[1]) * val(k[2])
        elif k[0] == "add":
            r = val(k[1]) + val(k[3]) if k[2] > 0 else val(k[1]) - val(k[3])
        elif k[0] == "rcp6":
            r = 1.0 / (6.0 * abs(val(k[1])))
        elif k[0] == "mulc":
            r = val(k[1]) * k[2]
        vals[i] = r
        return r
    v, bsum = [0.0] * 15, 0.0
    for guard, stmts in blocks:
        a, b = guard.split(" && ")
        if not (env[a] and env[b]):
            continue
        for slot, di, kab in stmts:
            x = kab[0] * val(kab[1])
            if di < 0 and at_lo:
                sc = scale(kab, 1000.0)
                bsum = bsum - (sc[0] * val(sc[1]))
            elif di > 0 and at_hi:
                sc = scale(kab, 100.0)
                bsum = bsum - (sc[0] * val(sc[1]))
            else:
                v[slot] += x
    return v, bsum


def main():
    blocks = build()

    # ---- emission ---------------------------------------------------------------------------------------------------
    uses = {}                                        # node id -> set of blocks using it (transitively)
    def walk(i, blk):
        if blk in uses.setdefault(i, set()):
            return
        uses[i].add(blk)
        k = defs[i]
        if k[0] == "mul": walk(k[1], blk); walk(k[2], blk)
        elif k[0] == "add": walk(k[1], blk); walk(k[3], blk)
        elif k[0] in ("rcp6", "mulc"): walk(k[1], blk)
    for bi, (_, stmts) in enumerate(blocks):
        for slot, di, kab in stmts:
            walk(kab[1], bi)
            if di != 0:                              # the Dirichlet branch multiplies by the boundary value
                walk(scale(kab, 1000.0 if di < 0 else 100.0)[1], bi)

    def text(i):
        k = defs[i]
        if k[0] == "sym": return k[1]
        return f"t{i}"
    def rhs(i):
        k = defs[i]
        if k[0] == "mul": return f"{text(k[1])} * {text(k[2])}"
        if k[0] == "add": return f"{text(k[1])} {'+' if k[2] > 0 else '-'} {text(k[3])}"
        if k[0] == "rcp6": return f"1.0 / (6.0 * fabs({text(k[1])}))"
        if k[0] == "mulc": return f"{text(k[1])} * {k[2]!r}"
        raise ValueError(k)
    emitted = set()
    out = []
    def emit(i, indent):
        k = defs[i]
        if k[0] == "sym" or i in emitted:
            return
        for dep in ([k[1], k[2]] if k[0] == "mul" else [k[1], k[3]] if k[0] == "add" else [k[1]]):
            emit(dep, indent)
        emitted.add(i)
        out.append(f"{indent}const double t{i} = {rhs(i)};")
    n_ops = {"mul": 0, "add": 0, "rcp6": 0, "mulc": 0}
    print("// GENERATED by tools/gen_kuhn_rows.py — do not edit.  The P1 row of a Kuhn-cube node with the arithmetic of the")
    print("// 24 incident tets partially evaluated (exact-zero operands eliminated, common subexpressions shared); every")
    print("// non-zero value is bit-identical to tet_G / K_ab evaluated in full.  See the generator for the identities used.")
    print("#pragma once")
    print("namespace heat {")
    print("// X[t], Y[t], Z[t]: coordinates of grid lines i-1+t, j-1+t, k-1+t; jlo/jhi/klo/khi: the cells at j-1 / j / k-1 / k exist;")
    print("// at_lo / at_hi: the x-neighbour at i-1 / i+1 is a Dirichlet node (1000 / 100).  v[15] and bsum must come in zeroed.")
    print("__device__ __forceinline__ void kuhn_row_p1_generated(const double (&X)[3], const double (&Y)[3], const double (&Z)[3], bool jlo, bool jhi,")
    print("                                                      bool klo, bool khi, bool at_lo, bool at_hi, double (&v)[15], double &bsum) {")
    shared = [i for i, b in uses.items() if len(b) > 1 and defs[i][0] != "sym"]
    for i in sorted(shared):
        emit(i, "    ")
    for l in out: print(l)
    for bi, (guard, stmts) in enumerate(blocks):
        out.clear()
        body = []
        for slot, di, kab in stmts:
            emit(kab[1], "        ")
            body.extend(out); out.clear()
            val = ("-" if kab[0] < 0 else "") + text(kab[1])
            if di == 0:
                body.append(f"        v[{slot}] += {val};")
            else:
                flag, bcv = ("at_lo", 1000.0) if di < 0 else ("at_hi", 100.0)
                sc = scale(kab, bcv)
                emit(sc[1], "        ")                 # at block level: the product is reused by later statements
                body.extend(out); out.clear()
                scv = ("-" if sc[0] < 0 else "") + text(sc[1])
                # bsum = bsum - t2 with t2 = kab * bc
                body.append(f"        if ({flag}) bsum = bsum - ({scv});")
                body.append(f"        else v[{slot}] += {val};")
        print(f"    if ({guard}) {{")
        for l in body: print(l)
        print("    }")
    print("}")
    print("}  // namespace heat")
    for i in emitted:
        n_ops[defs[i][0]] += 1
    print(f"// {len(emitted)} fp64 operations per interior row ({n_ops}), {sum(len(s) for _, s in blocks)} accumulations", file=sys.stderr)
    print(f"// {len(emitted)} fp64 temporaries: {n_ops}; {sum(len(s) for _, s in blocks)} accumulations into v / bsum")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Emit the hand-scheduled scatter-order sweep body (PX = 2, PY rows) as C++ statements.
src / dst: names of the value arrays (src[py][px] old, dst[py][px] new), w[py][px][f] weights (f: 0 W, 1 E, 2 S, 3 N),
fac[px], halo arrays hW[py], hE[py] (in registers at entry), hN[px], hS[px] (shared-memory loads issued at entry)."""
import sys

def gen(PY, src, dst, hw="hW", he="hE", nw="nW", ne="nE", next_shuffles=True, publish=True):
    L = []
    A = lambda i, j: "%s[%d][%d]" % (dst, i, j)
    X = lambda i, j: "%s[%d][%d]" % (src, i, j)
    W = lambda i, j, f: "w[%d][%d][%d]" % (i, j, f)
    def fma(acc, a, b): L.append("%s = fma(%s, %s, %s);" % (acc, a, b, acc))
    def mul(acc, a, b): L.append("%s = %s * %s;" % (acc, a, b))
    def HW(i): fma(A(i, 0), W(i, 0, 0), "%s[%d]" % (hw, i))
    def HE(i): fma(A(i, 1), W(i, 1, 1), "%s[%d]" % (he, i))
    def G0(i, s=True, rest=True):
        if s and i > 0: fma(A(i - 1, 0), W(i - 1, 0, 2), X(i, 0))
        if rest:
            if i + 1 < PY: mul(A(i + 1, 0), W(i + 1, 0, 3), X(i, 0))
            fma(A(i, 1), W(i, 1, 0), X(i, 0))
            fma(A(i, 0), "fac[0]", X(i, 0))
    def G1(i, s=True, rest=True):
        if s and i > 0: fma(A(i - 1, 1), W(i - 1, 1, 2), X(i, 1))
        if rest:
            if i + 1 < PY: mul(A(i + 1, 1), W(i + 1, 1, 3), X(i, 1))
            fma(A(i, 1), "fac[1]", X(i, 1))
            fma(A(i, 0), W(i, 0, 1), X(i, 1))
    def final(i, _unused=True):
        if next_shuffles:
            L.append("%s[%d] = __shfl_up_sync(0xffffffffu, %s, 1);" % (nw, i, A(i, 1)))
            L.append("%s[%d] = __shfl_down_sync(0xffffffffu, %s, 1);" % (ne, i, A(i, 0)))
        if publish and i == 0:
            L.append("pw[(0 * TH + r0) * PW + g] = %s; pw[(1 * TH + r0) * PW + g] = %s;" % (A(0, 0), A(0, 1)))
        if publish and i == PY - 1:
            L.append("pw[(0 * TH + r0 + %d) * PW + g] = %s; pw[(1 * TH + r0 + %d) * PW + g] = %s;" % (PY - 1, A(i, 0), PY - 1, A(i, 1)))
    # --- start: work that does not need the N halo (shared memory) ---
    mul(A(1, 0), W(1, 0, 3), X(0, 0))       # G0(0).N
    mul(A(1, 1), W(1, 1, 3), X(0, 1))       # G1(0).N
    HW(1)
    G0(1, s=False)
    G1(1, s=False)
    HE(1)
    # row 0 now that hN has arrived
    mul(A(0, 0), W(0, 0, 3), "hN[0]")
    mul(A(0, 1), W(0, 1, 3), "hN[1]")
    HW(0)
    # G0(0) rest without N (done), G1(0)
    fma(A(0, 1), W(0, 1, 0), X(0, 0))
    fma(A(0, 0), "fac[0]", X(0, 0))
    fma(A(0, 1), "fac[1]", X(0, 1))
    fma(A(0, 0), W(0, 0, 1), X(0, 1))
    HW(2)
    HE(0)
    # S products of row 1 into row 0
    fma(A(0, 0), W(0, 0, 2), X(1, 0))
    fma(A(0, 1), W(0, 1, 2), X(1, 1))
    # steady state: row i - 1 is final after the S products of row i; its shuffles / publish come a few instructions later
    pending = [0]
    for i in range(2, PY):
        if i > 2: HW(i)
        if i > 2: HE(i - 1)
        G0(i)
        for r in pending: final(r)
        pending = [i - 1]
        G1(i)
    HE(PY - 1)
    fma(A(PY - 1, 0), W(PY - 1, 0, 2), "hS[0]")
    fma(A(PY - 1, 1), W(PY - 1, 1, 2), "hS[1]")
    for r in pending: final(r)
    final(PY - 1, False)
    return L

if __name__ == "__main__":
    PY = int(sys.argv[1]); src = sys.argv[2]; dst = sys.argv[3]
    for l in gen(PY, src, dst): print("        " + l)

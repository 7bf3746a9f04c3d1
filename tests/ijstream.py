"""The SetValues / AddToValues call stream of `ref_dump -ijbuild MODE` (oracle/ref_dump.c: ijbuild), record for record,
and a plain-Python restatement of what the reference's IJ interface makes of such a stream
(IJ_mv/IJMatrix_parcsr.c:930-1000 insertion into the auxiliary rows, :2960-3045 assembly with the diagonal first,
:727-905 updates of existing entries after assembly).  Test infrastructure."""
import numpy as np


def calls_before_assembly(I, J, a, mode):
    """list of (add, rows, ncols, cols, vals) for phases 1-3; (I, J, a) = the generated operator"""
    N = I.size - 1
    out = []
    for r in range(N - 1, -1, -2):                                   # phase 1: SetValues, two rows per call
        rows, ncols, cols, vals = [], [], [], []
        for q in range(2):
            row = r - q
            if row < 0:
                break
            ln = int(I[row + 1] - I[row])
            rot = row % ln
            rows.append(row); ncols.append(ln)
            for k in range(ln):
                e = int(I[row]) + (k + rot) % ln
                cols.append(int(J[e])); vals.append(0.5 * float(a[e]))
        out.append((0, rows, ncols, cols, vals))
    for r in range(N):                                               # phase 2: AddToValues, one row per call
        ln = int(I[r + 1] - I[r])
        rot = (r + 1) % ln
        es = [int(I[r]) + (k + rot) % ln for k in range(ln)]
        out.append((1, [r], [ln], [int(J[e]) for e in es], [0.5 * float(a[e]) for e in es]))
    if mode == 2:
        for r in range(3, N, 7):                                     # phase 3: in-call duplicates, a row listed twice
            far, c2 = (r * 31 + 17) % N, (r * 13 + 5) % N
            mid = int(J[int(I[r]) + int(I[r + 1] - I[r]) // 2])
            out.append((0, [r], [3], [far, far, mid], [0.125, 0.25, 9.0]))
            out.append((1, [r, r], [1, 1], [c2, c2], [1.5, 2.5]))
    return out


def calls_after_assembly(EI, EJ, Ea):
    """the updates ref_dump applies to the assembled matrix (EI, EJ, Ea) before its second Assemble"""
    N = EI.size - 1
    out = [(1, [r], [1], [r], [1.0]) for r in range(0, N, 3)]
    for r in range(0, N, 5):
        if EI[r + 1] - EI[r] > 1:
            out.append((0, [r], [1], [int(EJ[EI[r] + 1])], [float(Ea[EI[r] + 1]) * 1.0]))
    return out


def replay(calls, N):
    """auxiliary rows after the calls: an entry is matched only against what its row held before the current
    (call, row) pair started; the first match is set / added to, anything else is appended"""
    rows = [[] for _ in range(N)]
    for add, rr, nc, cols, vals in calls:
        at = 0
        for row, n in zip(rr, nc):
            R = rows[row]
            old = len(R)
            for k in range(n):
                c, v = cols[at], vals[at]
                at += 1
                for q in range(old):
                    if R[q][0] == c:
                        R[q][1] = R[q][1] + v if add else v
                        break
                else:
                    R.append([c, v])
    return rows


def assemble(rows):
    """CSR with the LAST entry on the diagonal column first, the others in insertion order"""
    N = len(rows)
    I = np.zeros(N + 1, np.int32)
    J, A = [], []
    for i, R in enumerate(rows):
        dp = -1
        for q, (c, _) in enumerate(R):
            if c == i:
                dp = q
        if dp > -1:
            J.append(R[dp][0]); A.append(R[dp][1])
        for q, (c, v) in enumerate(R):
            if q != dp:
                J.append(c); A.append(v)
        I[i + 1] = len(J)
    return I, np.array(J, np.int32), np.array(A, np.float64)


def update(I, J, A, calls):
    """set / add on existing entries of the assembled matrix (first match from the start of the row); returns the
    number of records whose element does not exist"""
    missing = 0
    for add, rr, nc, cols, vals in calls:
        at = 0
        for row, n in zip(rr, nc):
            for k in range(n):
                c, v = cols[at], vals[at]
                at += 1
                for q in range(I[row], I[row + 1]):
                    if J[q] == c:
                        A[q] = A[q] + v if add else v
                        break
                else:
                    missing += 1
    return missing


def flat(call):
    add, rows, ncols, cols, vals = call
    return add, np.array(rows, np.int32), np.array(ncols, np.int32), np.array(cols, np.int32), np.array(vals, np.float64)

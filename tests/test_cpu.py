"""CPU suite (-m "not gpu"): the oracle against the reference-generated golden fixtures, the host
logic that needs no GPU, and that both shared libraries load and export every symbol their headers
declare.  No compute call reaches the CUDA library here."""
import glob
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from oracle import cport, refshim
from tests.util import V1, V2, golden_vec

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
SOLVES = json.load(open(os.path.join(GOLD, "solves.json")))


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "fields_*.npz")) + glob.glob(os.path.join(GOLD, "circ_*.npz"))))
def test_oracle_fields_match_reference_fixtures(path):
    G = np.load(path)
    npdim = tuple(int(v) for v in G["npdim"])
    m = cport.Mesh(npdim, grid=str(G["grid"]))
    for d, k in enumerate(("cx", "cy", "cz")):
        assert np.array_equal(m.coords[d], G[k])
    pde, vel, mu = str(G["pde"]), tuple(G["vel"]), float(G["mu"])
    A = cport.lhs(m, pde, vel, mu)
    assert np.array_equal(A, G["A"])
    for f in ("test_rhs", "manufactured_rhs", "manufactured_solution"):
        assert np.array_equal(cport.grid_function(m, pde, f, vel, mu), G[f])
    x, b = G["x"], G["manufactured_rhs"]
    assert np.array_equal(cport.apply(m, A, x), G["Ax"])
    assert np.array_equal(cport.apply_res(m, A, b, x), G["b_minus_Ax"])
    assert np.array_equal(cport.jacobi_apply(m, A, b), G["jacobi_b"])
    assert np.array_equal(cport.sgs_like_apply(m, A, cport.sgs_setup(m, A), b, 1), G["sgs_b"])
    assert np.array_equal(cport.sgs_like_apply(m, A, cport.strilu_setup(m, A, 1), b, 1), G["strilu_b"])
    y = b.copy(); cport.vecaxpy(m, 0.37, x, y)
    assert np.array_equal(y, G["b_plus_037x"])
    assert abs(cport.inner_vector_l2(m, x, b) - float(G["dot_x_b"])) <= 1e-13 * np.linalg.norm(x) * np.linalg.norm(b)
    assert abs(cport.norm_vector_l2(m, x) - float(G["norm_x"])) <= 1e-14 * float(G["norm_x"])
    assert abs(cport.norm_L2(m, x) - float(G["normL2_x"])) <= 1e-14 * float(G["normL2_x"])
    assert abs(cport.compute_error_L2(m, x, b) - float(G["errL2_x_b"])) <= 1e-14 * float(G["errL2_x_b"])
    assert m.h() == float(G["h"])


def _small(rec):
    return np.prod(rec["npdim"]) <= 48 ** 3 and rec["iters"] <= 1200


@pytest.mark.parametrize("rec", [r for r in SOLVES["solves"] if _small(r)], ids=lambda r: f"{r['case']}-{r['ksp']}-{r['pc']}")
def test_oracle_solvers_match_reference_runs(rec):
    m = cport.Mesh(rec["npdim"], grid=rec["grid"])
    A = cport.lhs(m, rec["pde"], rec["vel"], rec["mu"])
    b = golden_vec(m) if rec["rhs"] == "golden" else cport.grid_function(m, rec["pde"], rec["rhs"], rec["vel"], rec["mu"])
    x, info, hist = cport.solve(m, A, b, rec["ksp"], rec["pc"], 1e-6, 5000, 30, rec["apply_sweeps"], rec["build_sweeps"])
    tol_it = 0 if rec["ksp"] != "bcgs" else max(2, rec["iters"] // 20)   # BiCGSTAB amplifies rounding differences
    assert abs(info["iters"] - rec["iters"]) <= tol_it
    assert info["converged"] == rec["converged"]
    if rec["ksp"] != "bcgs" and rec["iters"] > 1:
        assert abs(info["resnorm"] - rec["resnorm"]) <= 1e-6 * rec["resnorm"]
        for step, val in rec["printed_rel_res"].items():
            assert abs(hist[int(step)] - val) <= 2e-5 * val      # %g prints 6 significant digits
        if rec["hist"]:
            h = np.array(rec["hist"])
            assert np.max(np.abs(hist - h) / h) <= 1e-6
    assert abs(cport.norm_vector_l2(m, b) - rec["bnorm"]) <= 1e-13 * rec["bnorm"]


def test_oracle_known_answers():
    for n, want in SOLVES["known"]["A_times_one_norm"].items():
        n = int(n)
        if n > 66:
            continue
        m = cport.Mesh((n, n, n))
        A = cport.lhs(m)
        x = m.zeros(); cport.vecset(m, 1.0, x)
        assert abs(cport.norm_vector_l2(m, cport.apply(m, A, x)) - want) <= 1e-13 * want
    for rec in SOLVES["known"]["matvectest"][:2]:
        n = rec["npdim"]
        m = cport.Mesh((n, n, n))
        A = cport.lhs(m)
        b, u = cport.grid_function(m, "poisson", "manufactured_rhs"), cport.grid_function(m, "poisson", "manufactured_solution")
        res = cport.apply_res(m, A, b, u)
        assert abs(cport.norm_vector_l2(m, res) / np.sqrt(float(n) ** 3) - rec["defect"]) <= 1e-12 * rec["defect"]


def test_redblack_sgs_is_the_same_operator_reordered():
    """M_rb = (D+L_rb) D^-1 (D+U_rb) applied by the oracle must solve M_rb z = r: check A-consistency via
    the identity (D+L)D^-1(D+U) z = r on the permuted ordering, using dense algebra on a tiny grid."""
    m = cport.Mesh((5, 6, 4), grid="chebyshev")
    A = cport.lhs(m, "convdiff", V1, 0.1)
    nx, ny, nz = m.n
    N = nx * ny * nz
    # dense A
    D = np.zeros((N, N))
    idx = lambda i, j, k: k * ny * nx + j * nx + i
    offs = [(-1, 0, 0), (0, -1, 0), (0, 0, -1), (0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1)]
    for k in range(nz):
        for j in range(ny):
            for i in range(nx):
                for s, (di, dj, dk) in enumerate(offs):
                    ii, jj, kk = i + di, j + dj, k + dk
                    if 0 <= ii < nx and 0 <= jj < ny and 0 <= kk < nz:
                        D[idx(i, j, k), idx(ii, jj, kk)] = A[s, k, j, i]
    colour = np.array([((i + 1) + (j + 1) + (k + 1)) & 1 for k in range(nz) for j in range(ny) for i in range(nx)])
    perm = np.concatenate([np.where(colour == 0)[0], np.where(colour == 1)[0]])
    P = D[np.ix_(perm, perm)]
    Dg = np.diag(np.diag(P)); L = np.tril(P, -1); U = np.triu(P, 1)
    Mrb = (Dg + L) @ np.linalg.inv(Dg) @ (Dg + U)
    rng = np.random.default_rng(0)
    r = m.zeros(); r[1:-1, 1:-1, 1:-1] = rng.uniform(-1, 1, (nz, ny, nx))
    z = cport.sgs_redblack_apply(m, A, cport.sgs_setup(m, A), r)
    zz = z[1:-1, 1:-1, 1:-1].ravel()[perm]
    rr = r[1:-1, 1:-1, 1:-1].ravel()[perm]
    assert np.linalg.norm(Mrb @ zz - rr) <= 1e-12 * np.linalg.norm(rr)
    # and the lexicographic one against the unpermuted splitting
    Dg = np.diag(np.diag(D)); L = np.tril(D, -1); U = np.triu(D, 1)
    M = (Dg + L) @ np.linalg.inv(Dg) @ (Dg + U)
    z = cport.sgs_like_apply(m, A, cport.sgs_setup(m, A), r, 1)
    assert np.linalg.norm(M @ z[1:-1, 1:-1, 1:-1].ravel() - r[1:-1, 1:-1, 1:-1].ravel()) <= 1e-12 * np.linalg.norm(rr)


def _sgw_row(P, fwd, b, c):
    """Python statement of sgw_row<P, FWD> in struct3d_b200/csrc/sgs.cu (the row a (b, c) pencil's inputs occupy in shared memory)."""
    if P == 2:
        lo, hi = 5 * (b >> 1) + 2 * (c >> 1), (b >> 2) * 4 + (c & 1) * 2 + (b & 1)
    else:
        lo, hi = ((b + c + 1) >> 1) + (c >> 1) + 4 * (b & 1), 2 * (b & 1) + (c & 1)
    return 8 * hi + ((lo if fwd else -lo) & 7)


@pytest.mark.parametrize("P", [1, 2])
@pytest.mark.parametrize("fwd", [True, False])
@pytest.mark.parametrize("TI", [16, 32])
def test_sgs_shared_memory_row_map_is_conflict_free(P, fwd, TI):
    """The warp-per-tile SGS kernel addresses a tile's inputs as row(pencil) * (TI + 2) + i, and lane (m, c) reads i = h - s
    (forward) or ti - 1 - (h - s) (backward) at step h, s = b + c.  The map must (1) be a bijection of the 32 P pencils onto rows,
    (2) put a lane's second pencil 8 rows after its first (one shared pointer + immediates), (3) give the 16 lanes of every
    half-warp 16 different 8-byte banks at every step, for both pencils."""
    R = TI + 2
    rows = {_sgw_row(P, fwd, b, c) for b in range(4 * P) for c in range(8)}
    assert rows == set(range(32 * P))
    for lane in range(32):
        m, c = lane >> 3, lane & 7
        if P == 2:
            assert _sgw_row(2, fwd, 2 * m + 1, c) == _sgw_row(2, fwd, 2 * m, c) + 8
    for e in range(P):
        for h in range(TI + 4 * P + 8):
            for half in (range(0, 16), range(16, 32)):
                banks = set()
                for lane in half:
                    m, c = lane >> 3, lane & 7
                    b = P * m + e
                    u = h - (b + c)
                    i = u if fwd else TI - 1 - u
                    banks.add((_sgw_row(P, fwd, b, c) * R + i) % 16)
                assert len(banks) == 16, (P, fwd, TI, e, h)


@pytest.mark.parametrize("nranks", [2, 3, 4])
def test_peer_memory_halo_protocol_needs_no_acknowledgement(nranks):
    """Model of the peer-memory halo exchange (struct3d_b200/csrc/context.cu): every rank runs, in stream order, push(1), consume(1),
    push(2), consume(2), ...  push(e) writes this rank's boundary planes into the neighbours' window slots of parity e & 1 and then
    their flags; consume(e) (the pull kernel, or the boundary blocks of the SpMV) cannot finish before BOTH neighbours' push(e) have.
    Claim: two slots per direction suffice without acknowledgements -- whenever a rank overwrites a neighbour's slot with epoch e,
    that neighbour has already consumed epoch e - 2.  Checked over every reachable interleaving of the ranks."""
    nepochs = 5
    nops = 2 * nepochs                        # op index 2(e-1) = push(e), 2(e-1)+1 = consume(e)
    start = tuple([0] * nranks)
    seen, todo = {start}, [start]
    while todo:
        pc = todo.pop()
        for r in range(nranks):
            if pc[r] >= nops:
                continue
            e, is_consume = pc[r] // 2 + 1, pc[r] % 2 == 1
            nbrs = [q for q in (r - 1, r + 1) if 0 <= q < nranks]
            if is_consume:
                if not all(pc[q] > 2 * (e - 1) for q in nbrs):      # a neighbour has not pushed epoch e yet: the kernel waits
                    continue
            else:
                for q in nbrs:                                       # push(e) overwrites the neighbour's slot of parity e & 1
                    assert e <= 2 or pc[q] > 2 * (e - 3) + 1, f"rank {r} overwrites epoch {e - 2} in rank {q}'s window before it was consumed"
            nxt = pc[:r] + (pc[r] + 1,) + pc[r + 1:]
            if nxt not in seen:
                seen.add(nxt)
                todo.append(nxt)
    assert tuple([nops] * nranks) in seen                            # and the schedule cannot deadlock


@pytest.mark.parametrize("nranks", [2, 3, 4])
def test_peer_memory_allreduce_needs_no_acknowledgement(nranks):
    """Model of the one-shot all-reduce over peer memory (allreduce_peer_kernel, struct3d_b200/csrc/context.cu): the kernels of a rank
    run in stream order, reduce(1), reduce(2), ...; reduce(e) first WRITES this rank's values and then the epoch into slot [e & 1][rank]
    of EVERY rank's window, then waits until all slots [e & 1][*] of its OWN window carry epoch e, then READS them.  Modelled as two
    events per reduction (write, read-done), every interleaving explored.  Claims: (1) when a rank overwrites a slot with epoch e, the
    owner of that window has finished reading epoch e - 2 from it -- two parities suffice, no acknowledgement; (2) no deadlock; (3) what a
    rank reads for epoch e is what every rank wrote for epoch e."""
    nepochs = 5
    nops = 2 * nepochs                        # op 2(e-1) = write(e) into all windows, op 2(e-1)+1 = read(e) from the own window
    start = tuple([0] * nranks)
    seen, todo = {start}, [start]
    while todo:
        pc = todo.pop()
        for r in range(nranks):
            if pc[r] >= nops:
                continue
            e, is_read = pc[r] // 2 + 1, pc[r] % 2 == 1
            if is_read:
                if not all(pc[q] > 2 * (e - 1) for q in range(nranks)):   # somebody's epoch e has not landed: the kernel spins
                    continue
                # what sits in slot [e & 1][q] of r's window is q's LATEST write of that parity: it must be epoch e, not e + 2
                for q in range(nranks):
                    latest = (pc[q] + 1) // 2                              # epochs q has written so far
                    newest_same_parity = latest if latest % 2 == e % 2 else latest - 1
                    assert newest_same_parity == e, f"rank {r} reads epoch {e} of rank {q} after it was overwritten by {newest_same_parity}"
            else:
                for q in range(nranks):                                    # write(e) overwrites slot [e & 1][r] in q's window
                    assert e <= 2 or pc[q] > 2 * (e - 3) + 1, f"rank {r} overwrites epoch {e - 2} in rank {q}'s window before it was read"
            nxt = pc[:r] + (pc[r] + 1,) + pc[r + 1:]
            if nxt not in seen:
                seen.add(nxt)
                todo.append(nxt)
    assert tuple([nops] * nranks) in seen


@pytest.mark.parametrize("dims", [(2, 2, 1), (3, 2, 2), (2, 3, 2)])
def test_sgs_tile_mailboxes_are_clean_after_every_sweep(dims):
    """Model of the mailbox hand-off of the warp-per-tile SGS kernel (sgs_warpflow_kernel, struct3d_b200/csrc/sgs.cu): every tile owns
    three boxes (its last-i column, last-j rows, last-k rows) that hold a sentinel outside a sweep.  A tile polls the boxes of its
    predecessors (in sweep order) until they are full, re-arms them, sweeps, and fills its own boxes -- but ONLY those a tile of this
    sweep will read.  Claims, over forward and backward sweeps alternating and EVERY order in which ready tiles can run: a consumer never
    sees a box written by an earlier sweep (stale faces), a producer never overwrites a full box, and all boxes are empty again when a
    sweep ends (so the next sweep, whose producers and consumers swap roles, starts clean)."""
    import itertools
    nx, ny, nz = dims
    tiles = list(itertools.product(range(nx), range(ny), range(nz)))
    boxes = {(t, d): None for t in tiles for d in range(3)}        # (tile, direction) -> sweep number that filled it, None = sentinel

    def neighbour(t, d, step):
        q = list(t); q[d] += step
        return tuple(q) if 0 <= q[d] < dims[d] else None

    def run_sweep(n, fwd, order_seed):
        step = 1 if fwd else -1
        done = set()
        rng = np.random.default_rng(order_seed)
        while len(done) < len(tiles):
            ready = [t for t in tiles if t not in done and all(neighbour(t, d, -step) is None or boxes[(neighbour(t, d, -step), d)] is not None for d in range(3))]
            assert ready, "deadlock: no tile can run"
            t = ready[rng.integers(len(ready))]
            for d in range(3):
                p = neighbour(t, d, -step)
                if p is not None:
                    assert boxes[(p, d)] == n, f"tile {t} takes a face of sweep {boxes[(p, d)]} in sweep {n}"
                    boxes[(p, d)] = None                           # re-armed by its reader
            for d in range(3):
                if neighbour(t, d, step) is not None:              # only where a tile of this sweep will read it
                    assert boxes[(t, d)] is None, f"tile {t} overwrites a full box in sweep {n}"
                    boxes[(t, d)] = n
            done.add(t)
        assert all(v is None for v in boxes.values()), f"boxes left full after sweep {n}"

    for seed in range(40):
        for n in range(4):
            run_sweep(n, fwd=(n % 2 == 0), order_seed=1000 * seed + n)


@pytest.mark.parametrize("nranks", [2, 3, 4])
def test_sgs_slab_window_is_safe_with_one_buffer_per_direction(nranks):
    """Model of the exact SGS sweep through z-slabs (sgs.cu / context.cu): every rank runs forward sweep, backward sweep, forward, ...
    as kernels in stream order.  While rank p runs a forward sweep it WRITES k-face rows into rank p+1's forward window and READS its
    own forward window (filled by rank p-1); it can only finish after rank p-1 has finished the same sweep (its last plane of tiles
    comes last).  Backward sweeps mirror that.  Claim: one window plane per direction suffices -- when a rank starts writing sweep n
    into a neighbour's window, that neighbour has finished reading sweep n-1 of the same direction.  Every interleaving is explored."""
    napplies = 3
    nev = 4 * napplies                         # per rank: start F1, end F1, start B1, end B1, start F2, ...
    start = tuple([0] * nranks)
    seen, todo = {start}, [start]

    def ended(pc_q, k, backward):              # has rank q finished sweep (k, direction)?
        return pc_q > 4 * (k - 1) + (3 if backward else 1)

    while todo:
        pc = todo.pop()
        for r in range(nranks):
            if pc[r] >= nev:
                continue
            k, phase = pc[r] // 4 + 1, pc[r] % 4          # phase 0 start F, 1 end F, 2 start B, 3 end B
            backward, is_end = phase >= 2, phase % 2 == 1
            src = r + 1 if backward else r - 1             # whose rows this rank reads
            dst = r - 1 if backward else r + 1             # whose window this rank writes
            if is_end:
                if 0 <= src < nranks and not ended(pc[src], k, backward):
                    continue                                # still waiting for the neighbour's last plane of tiles
            elif 0 <= dst < nranks and k > 1:
                assert ended(pc[dst], k - 1, backward), f"rank {r} starts sweep {k} into rank {dst}'s window while it still reads sweep {k - 1}"
            nxt = pc[:r] + (pc[r] + 1,) + pc[r + 1:]
            if nxt not in seen:
                seen.add(nxt)
                todo.append(nxt)
    assert tuple([nev] * nranks) in seen


@pytest.mark.parametrize("dims,nworkers", [((2, 2, 2), 1), ((2, 2, 2), 3), ((3, 2, 2), 2), ((2, 3, 3), 4)])
def test_sgs_tile_dispatch_cannot_deadlock(dims, nworkers):
    """Model of the dataflow dispatch of the exact SGS sweep (sgs.cu): tiles are handed out in hyperplane order (ci + cj + ck, then
    lexicographic); worker w starts with ticket w, and claims its NEXT ticket from a shared counter as soon as the predecessors of
    its current tile are done (the kernel prefetches the ticket before it sweeps); a tile may be swept only after its three
    predecessors.  Every reachable interleaving must be able to finish all tiles, with any number of workers."""
    ntx, nty, ntz = dims
    tiles = sorted(((ci, cj, ck) for ck in range(ntz) for cj in range(nty) for ci in range(ntx)), key=lambda t: (sum(t), t[2], t[1], t[0]))
    index = {t: n for n, t in enumerate(tiles)}
    preds = [[index[p] for p in ((ci - 1, cj, ck), (ci, cj - 1, ck), (ci, cj, ck - 1)) if p in index] for ci, cj, ck in tiles]
    nt = len(tiles)
    # state: (frozenset of finished tiles, per worker (current ticket or None, prefetched ticket or None), next counter value)
    init = (frozenset(), tuple((w if w < nt else None, None) for w in range(nworkers)), nworkers)
    seen, todo, finished = {init}, [init], False
    while todo:
        done, workers, counter = todo.pop()
        if len(done) == nt:
            finished = True
            continue
        moves = 0
        for w, (cur, nxt) in enumerate(workers):
            if cur is None:
                continue
            if not all(p in done for p in preds[cur]):
                continue                                              # spinning on a flag
            moves += 1
            if nxt is None and counter < nt + nworkers:               # after the wait: claim the next ticket, then sweep
                st = (done, workers[:w] + ((cur, counter),) + workers[w + 1:], counter + 1)
            else:                                                     # sweep, publish, move on to the prefetched ticket
                st = (done | {cur}, workers[:w] + ((nxt if (nxt is not None and nxt < nt) else None, None),) + workers[w + 1:], counter)
            if st not in seen:
                seen.add(st)
                todo.append(st)
        assert moves > 0, f"deadlock with finished={sorted(done)} workers={workers}"
    assert finished


def test_jacobi_sweeps_statement_converges_to_the_exact_sweep():
    """orc_sgs_jacobi_sweeps_apply is not in the reference (it is the deterministic counterpart of its asynchronous threaded
    sweeps), so pin it by its defining properties: n synchronous Jacobi sweeps on a triangular system with
    dependency depth d = nx+ny+nz-2 are exact from n = d on (bit-identical to the sequential sweep), monotonically closer before,
    and one sweep from zero is the Jacobi-scaled right-hand side."""
    m = cport.Mesh((9, 7, 8), grid="chebyshev")
    A = cport.lhs(m, "convdiff", V1, 0.1)
    di = cport.sgs_setup(m, A)
    rng = np.random.default_rng(3)
    r = m.zeros(); r[1:-1, 1:-1, 1:-1] = rng.uniform(-1, 1, tuple(reversed(m.n)))
    exact = cport.sgs_like_apply(m, A, di, r, 1)
    depth = sum(m.n) - 2
    assert np.array_equal(cport.sgs_jacobi_sweeps_apply(m, A, di, r, depth), exact)
    assert np.array_equal(cport.sgs_jacobi_sweeps_apply(m, A, di, r, depth + 3), exact)
    errs = [np.abs(cport.sgs_jacobi_sweeps_apply(m, A, di, r, n) - exact).max() for n in (1, 2, 4, 8, depth - 1)]
    assert errs[0] > errs[2] > errs[3] and errs[-1] <= errs[3]
    assert errs[-1] > 0.0 or depth <= 2        # one sweep short of the depth is not yet exact (generic data)
    one = cport.sgs_jacobi_sweeps_apply(m, A, di, r, 1)
    # forward sweep 1 from zero: y = r * dinv; backward sweep 1 from zero: z = y  =>  z = r / diag
    assert np.array_equal(one[1:-1, 1:-1, 1:-1], (r[1:-1, 1:-1, 1:-1] * di))


@pytest.mark.skipif(not refshim.available(), reason="oracle/_ref not built (no /root/reference here)")
def test_oracle_convdiff_circular_matches_live_reference():
    """the rotational-velocity operator, reproduced as the reference has it."""
    refshim.set_num_threads(1)
    for npdim, grid in (((14, 11, 9), "chebyshev"), ((12, 12, 12), "uniform")):
        m, rm = cport.Mesh(npdim, grid=grid), refshim.Mesh(npdim, grid=grid)
        P = refshim.Pde("convdiff_circular", (1.3, 0, 0), 0.07)
        assert np.array_equal(cport.lhs(m, "convdiff_circular", (1.3, 0, 0), 0.07), P.lhs(rm).stacked())
        for w in ("test_rhs", "manufactured_solution", "manufactured_rhs"):
            assert np.array_equal(cport.grid_function(m, "convdiff_circular", w, (1.3, 0, 0), 0.07), P.vector(rm, w).a)


@pytest.mark.skipif(not refshim.available(), reason="oracle/_ref not built (no /root/reference here)")
def test_oracle_matches_live_reference():
    """Where the compiled reference is present, re-pin the restatement against it on a fresh case."""
    refshim.set_num_threads(1)
    npdim = (14, 11, 9)
    m, rm = cport.Mesh(npdim, grid="chebyshev"), refshim.Mesh(npdim, grid="chebyshev")
    P = refshim.Pde("convdiff", V2, 0.07)
    RA = P.lhs(rm)
    A = cport.lhs(m, "convdiff", V2, 0.07)
    assert np.array_equal(A, RA.stacked())
    b = P.vector(rm, "manufactured_rhs")
    y = refshim.Vec(rm)
    refshim.apply(RA, b, y)
    assert np.array_equal(cport.apply(m, A, b.a), y.a)
    x = refshim.Vec(rm)
    info = refshim.solve(RA, b, x, {"-s3d_ksp_type": "gcr", "-s3d_pc_type": "sgs", "-ksp_rtol": 1e-8, "-ksp_max_it": 500,
                                   "-s3d_pc_apply_sweeps": 1, "-s3d_thread_chunk_size": 64, "-s3d_pc_use_threaded_apply": "false"})
    xo, oi, _ = cport.solve(m, A, b.a, "gcr", "sgs", 1e-8, 500)
    assert oi["iters"] == info["iters"] and np.abs(xo - x.a).max() <= 1e-12


# ------------------------------------------------------------------ libraries and host logic (no GPU)

def _declared(header, prefix):
    txt = open(os.path.join(ROOT, "include", header)).read()
    return sorted(set(re.findall(r"\b(" + prefix + r"[A-Za-z0-9_]+)\s*\(", txt)))


def test_cuda_library_exports_every_declared_symbol():
    from struct3d_b200 import capi
    L = capi.lib()
    decl = _declared("s3d_cuda.h", "s3d_")
    assert len(decl) >= 55
    for name in decl:
        assert hasattr(L, name), name
    assert sorted(capi.EXPORTS) == decl


def test_host_library_exports_every_declared_symbol():
    from struct3d_b200 import hostapi
    L = hostapi.lib()
    decl = _declared("s3d_host.h", "s3dh_")
    for name in decl:
        assert hasattr(L, name), name
    assert sorted(hostapi.EXPORTS) == decl


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from struct3d_b200 import capi, hostapi
    with pytest.raises(capi.S3dError, match="no CPU fallback"):
        capi.Context(0)
    assert hostapi.lib().s3dh_init(-1, 0, 1, None) == -1
    assert b"no CPU fallback" in hostapi.lib().s3dh_last_error()


def test_grid_layout_and_slab_decomposition():
    from struct3d_b200 import capi
    g = capi.make_grid(512, 512, 512)
    assert (g.vpitch % 16, g.cpitch % 16, g.voff % 16, g.vlead) == (0, 0, 0, 16)
    assert g.vpitch >= g.vlead + 512 + 1 and g.vplane == g.vpitch * 514 and g.vlen == g.vplane * 514
    assert g.cplane == g.cpitch * 512 and g.cstride >= g.cplane * 512
    for nz, P in ((1024, 8), (37, 4), (5, 5), (130, 3)):
        covered = 0
        for r in range(P):
            s = capi.make_grid(64, 48, nz, r, P)
            assert s.k0 == covered and s.nz >= 1 and (s.nx, s.ny, s.nz_global, s.rank, s.nranks) == (64, 48, nz, r, P)
            covered += s.nz
        assert covered == nz
    with pytest.raises(ValueError):
        capi.make_grid(8, 8, 3, 0, 4)      # fewer planes than ranks
    with pytest.raises(ValueError):
        capi.make_grid(0, 8, 8)


def test_control_file_reader_matches_reference_format():
    from struct3d_b200 import hostapi
    c = hostapi.read_control(os.path.join(ROOT, "tests", "input", "convdiff64.control"))
    assert c["npdim"] == [64, 64, 64] and c["grid"] == "uniform" and c["pde"] == "convdiff" and c["nruns"] == 2
    assert c["vel"] == [0.577350269189626, 0.7, 0.3] and c["mu"] == 0.1 and c["btypes"] == [1] * 6 and c["bvals"] == [0.0] * 6
    c = hostapi.read_control(os.path.join(ROOT, "tests", "input", "poisson16_cheb.control"))
    assert c["grid"] == "chebyshev" and c["pde"] == "poisson" and c["rmin"] == [-1.0] * 3 and c["rmax"] == [1.0] * 3
    if refshim.available():   # the reference's own reader parses our files identically
        for f in ("convdiff64", "poisson16_cheb", "poisson26", "convdiff48"):
            p = os.path.join(ROOT, "tests", "input", f + ".control")
            assert refshim.read_control(p) == hostapi.read_control(p)
    with pytest.raises(hostapi.HostError):
        hostapi.read_control("/nonexistent.control")


def test_options_store(tmp_path):
    from struct3d_b200 import hostapi
    hostapi.lib().s3dh_options_clear()
    hostapi.load_options(os.path.join(ROOT, "tests", "input", "gcr_jacobi.perc"))
    # creating a solver needs a GPU; the store itself is host logic: check through a second load overriding a key
    p = tmp_path / "o.perc"
    p.write_text("# comment\n-ksp_rtol 1e-9\n-flag_without_value\n-s3d_ksp_type cg # trailing\n")
    hostapi.load_options(str(p))
    with pytest.raises(hostapi.HostError):
        hostapi.load_options(str(tmp_path / "missing.perc"))


WORKER = r"""
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, os.environ["S3D_ROOT"])
from oracle import cport
from struct3d_b200 import capi
from tests.util import golden_vec

dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + os.environ["S3D_PORT"], rank=int(os.environ["RANK"]), world_size=2)
rank, P = dist.get_rank(), 2
npdim = (12, 10, 15)
m = cport.Mesh(npdim, grid="chebyshev")
A = cport.lhs(m, "convdiff", (0.3, -0.7, 0.5), 0.1)
x = golden_vec(m)
y_global = cport.apply(m, A, x)
nx, ny, nz = m.n
g = capi.make_grid(nx, ny, nz, rank, P)                       # the library's own slab arithmetic
k0, nzl = g.k0, g.nz
# local slab problem: coordinates along z restricted to [k0, k0+nzl+2)
lm = cport.Mesh((npdim[0], npdim[1], nzl + 2), grid="chebyshev")
lm.coords[2] = m.coords[2][k0:k0 + nzl + 2].copy()
lA = np.ascontiguousarray(A[:, k0:k0 + nzl])
lx = np.ascontiguousarray(x[k0:k0 + nzl + 2]).copy()
lx[0] = 0; lx[-1] = 0                                          # ghosts start empty
# halo exchange exactly as s3d_halo_exchange does: top real plane up, bottom real plane down
import torch
reqs = []
if rank + 1 < P:
    reqs.append(dist.isend(torch.from_numpy(lx[nzl].copy()), rank + 1))
    up = torch.empty(lx[0].shape, dtype=torch.float64); reqs.append(dist.irecv(up, rank + 1))
if rank - 1 >= 0:
    reqs.append(dist.isend(torch.from_numpy(lx[1].copy()), rank - 1))
    dn = torch.empty(lx[0].shape, dtype=torch.float64); reqs.append(dist.irecv(dn, rank - 1))
for r in reqs: r.wait()
if rank + 1 < P: lx[nzl + 1] = up.numpy()
if rank - 1 >= 0: lx[0] = dn.numpy()
ly = cport.apply(lm, lA, lx)
assert np.array_equal(ly[1:-1], y_global[k0 + 1:k0 + nzl + 1]), "slab SpMV differs from the global one"
# dot product = allreduce of slab partial sums
part = torch.tensor([cport.inner_vector_l2(lm, ly, ly)], dtype=torch.float64)
dist.all_reduce(part)
full = cport.inner_vector_l2(m, y_global, y_global)
assert abs(part.item() - full) <= 1e-13 * full
# the exact SGS sweep THROUGH the slabs (sgs.cu + the peer-memory window of context.cu, stated on the CPU): forward sweeps run rank 0
# then rank 1 -- rank 1's bottom ghost plane holds rank 0's top plane of y --, backward sweeps run rank 1 then rank 0; the pieces
# must be the sequential sweep over the whole grid, bit for bit
dinv = cport.sgs_setup(m, A)
rr = golden_vec(m)[::-1].copy(); rr[0] = 0; rr[-1] = 0
z_global = cport.sgs_like_apply(m, A, dinv, rr, 1)
ldinv = np.ascontiguousarray(dinv[k0:k0 + nzl])
lr = np.ascontiguousarray(rr[k0:k0 + nzl + 2]).copy()
ly2, lz2 = np.zeros_like(lr), np.zeros_like(lr)
if rank > 0:                                                     # wait for the slab below: its last plane of y
    t = torch.empty(lr[0].shape, dtype=torch.float64); dist.recv(t, rank - 1); ly2[0] = t.numpy()
cport.sgs_forward_inplace(lm, lA, ldinv, lr, ly2)
if rank + 1 < P:
    dist.send(torch.from_numpy(ly2[nzl].copy()), rank + 1)
    t = torch.empty(lr[0].shape, dtype=torch.float64); dist.recv(t, rank + 1); lz2[nzl + 1] = t.numpy()   # the slab above: its first plane of z
cport.sgs_backward_inplace(lm, lA, ldinv, ly2, lz2)
if rank > 0:
    dist.send(torch.from_numpy(lz2[1].copy()), rank - 1)
assert np.array_equal(lz2[1:-1], z_global[k0 + 1:k0 + nzl + 1]), "SGS through the slabs differs from the sequential sweep over the whole grid"
# the OVERLAPPING ordering (S3D_SGS_OVERLAP, the default on slabs), stated through the ranks: what s3d_sgs_overlap_setup does once per
# operator (plane nz-1 of every coefficient array and of dinv into plane -1 of the rank above, plane 0 into plane nz of the rank below)
# and what s3d_sgs_apply does per application (one-plane halo exchange of r, then the sequential sweep over planes -1 .. nz with zero
# inflow beyond, own planes kept); no rank waits for another one's SWEEP.  Must be the oracle's whole-grid statement, bit for bit.
def swap_planes(lo_plane, hi_plane):
    # send my bottom plane down and my top plane up; returns (plane from below, plane from above), None at the ends
    got = [None, None]
    rq = []
    if rank + 1 < P:
        rq.append(dist.isend(torch.from_numpy(np.ascontiguousarray(hi_plane)), rank + 1))
        got[1] = torch.empty(hi_plane.shape, dtype=torch.float64); rq.append(dist.irecv(got[1], rank + 1))
    if rank - 1 >= 0:
        rq.append(dist.isend(torch.from_numpy(np.ascontiguousarray(lo_plane)), rank - 1))
        got[0] = torch.empty(lo_plane.shape, dtype=torch.float64); rq.append(dist.irecv(got[0], rank - 1))
    for q in rq: q.wait()
    return [None if t is None else t.numpy() for t in got]
z_ras = cport.sgs_overlap_apply(m, A, dinv, rr, P, 1)
assert not np.array_equal(z_ras, cport.sgs_overlap_apply(m, A, dinv, rr, P, 0))
A_lo, A_hi = swap_planes(lA[:, 0], lA[:, nzl - 1])              # set-up: coefficient planes
d_lo, d_hi = swap_planes(ldinv[0], ldinv[nzl - 1])
lr3 = np.ascontiguousarray(rr[k0:k0 + nzl + 2]).copy(); lr3[0] = 0; lr3[-1] = 0
r_lo, r_hi = swap_planes(lr3[1], lr3[nzl])                      # per application: the halo exchange of r
if r_lo is not None: lr3[0] = r_lo
if r_hi is not None: lr3[nzl + 1] = r_hi
lo, hi = int(rank > 0), int(rank + 1 < P)
xA = np.concatenate(([A_lo[:, None]] if lo else []) + [lA] + ([A_hi[:, None]] if hi else []), axis=1)
xd = np.concatenate(([d_lo[None]] if lo else []) + [ldinv] + ([d_hi[None]] if hi else []), axis=0)
xm = cport.Mesh((npdim[0], npdim[1], nzl + lo + hi + 2))
xr = xm.zeros(); xr[1:-1] = lr3[1 - lo:nzl + 1 + hi]
xz = cport.sgs_like_apply(xm, np.ascontiguousarray(xA), np.ascontiguousarray(xd), xr, 1)
assert np.array_equal(xz[1 + lo:1 + lo + nzl], z_ras[k0 + 1:k0 + nzl + 1]), "overlapping SGS through the ranks differs from the whole-grid statement"
# NCCL id plumbing used by bench.py: rank 0's bytes reach every rank unchanged
obj = [bytes(range(128)) if rank == 0 else None]
dist.broadcast_object_list(obj, src=0)
assert obj[0] == bytes(range(128))
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_overlapping_slab_sgs_meets_the_iteration_gate():
    """north_star's gate for a parallel SGS (the same residual within 5 % more iterations than the reference's sequential SGS), for the
    ordering that is the default on z-slabs, on the CPU with the oracle's statement of it (cport.sgs_overlap_apply), convection-diffusion:
    (1) Richardson -- a stationary iteration, so its count measures the preconditioner itself, smoothly and deterministically: within
    5 % on 2 / 4 / 8 slabs (down to 8 planes per slab), where the slab-local ordering (no overlap) needs +15 %;
    (2) BiCGSTAB (BASELINE's solver; its counts move by several per cent with anything) on a larger grid, 2 and 4 slabs."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import sgs_overlap_study as st
    vel, mu = (0.577350269189626, 0.7, 0.3), 0.1

    def richardson(m, A, b, pc, rtol=1e-6, maxit=2000):
        x, bb = m.zeros(), np.sqrt(np.vdot(b, b))
        for it in range(maxit):
            r = cport.apply_res(m, A, b, x)
            if np.sqrt(np.vdot(r, r)) / bb <= rtol:
                return it
            x += pc(r)
        return maxit

    m = cport.Mesh((50, 50, 66))
    A = cport.lhs(m, "convdiff", vel, mu)
    b = cport.grid_function(m, "convdiff", "manufactured_rhs", vel, mu)
    dinv = cport.sgs_setup(m, A)
    exact = richardson(m, A, b, lambda r: cport.sgs_like_apply(m, A, dinv, r, 1))
    assert 250 <= exact <= 400
    for nslabs in (2, 4, 8):
        it = richardson(m, A, b, lambda r: cport.sgs_overlap_apply(m, A, dinv, r, nslabs, 1))
        assert it <= 1.05 * exact, f"Richardson, {nslabs} overlapping slabs: {it} iterations against {exact}"
    local8 = richardson(m, A, b, lambda r: cport.sgs_overlap_apply(m, A, dinv, r, 8, 0))
    assert local8 > 1.10 * exact, (local8, exact)

    m = cport.Mesh((66, 66, 130))
    A = cport.lhs(m, "convdiff", vel, mu)
    b = cport.grid_function(m, "convdiff", "manufactured_rhs", vel, mu)
    dinv = cport.sgs_setup(m, A)
    exact = st.bicgstab(m, A, b, lambda r: cport.sgs_like_apply(m, A, dinv, r, 1))
    for nslabs in (2, 4):
        it = st.bicgstab(m, A, b, lambda r: cport.sgs_overlap_apply(m, A, dinv, r, nslabs, 1))
        assert it <= 1.05 * exact, f"BiCGSTAB, {nslabs} overlapping slabs: {it} iterations against {exact}"


def test_blocks_ordering_statement_and_gate():
    """The oracle's statement of the blocks ordering (S3D_SGS_BLOCKS): the block partition covers the axis with blocks of 7 modulo 8 planes (the
    last one takes the rest, at least 8), one block is the reference's sequential sweep bit for bit, and Richardson on convection-diffusion
    stays within north_star's +5 % of the sequential sweep's iterations for the block sizes the library picks by default (about 32)."""
    for n in range(8, 140):
        for B in (1, 2, 3, 4, 8, 16):
            b = cport.block_bounds(n, B)
            assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(len(b) - 1)) and len(b) <= max(1, min(B, n // 8))
            assert all((k1 - k0) % 8 == 7 for k0, k1 in b[:-1]) and (len(b) == 1 or b[-1][1] - b[-1][0] >= 8)
    vel, mu = (0.577350269189626, 0.7, 0.3), 0.1
    m = cport.Mesh((34, 66, 66))
    A = cport.lhs(m, "convdiff", vel, mu)
    b = cport.grid_function(m, "convdiff", "manufactured_rhs", vel, mu)
    dinv = cport.sgs_setup(m, A)
    r = golden_vec(m)
    assert np.array_equal(cport.sgs_blocks_apply(m, A, dinv, r, 1, 1), cport.sgs_like_apply(m, A, dinv, r, 1))
    assert cport.blocks_by_size(*m.n) == (2, 2)

    def richardson(pc, rtol=1e-6, maxit=3000):
        x, bb = m.zeros(), np.sqrt(np.vdot(b, b))
        for it in range(maxit):
            res = cport.apply_res(m, A, b, x)
            if np.sqrt(np.vdot(res, res)) / bb <= rtol:
                return it
            x += pc(res)
        return maxit

    exact = richardson(lambda v: cport.sgs_like_apply(m, A, dinv, v, 1))
    blocks = richardson(lambda v: cport.sgs_blocks_apply(m, A, dinv, v, 2, 2))
    assert 100 <= exact <= 1000 and blocks <= 1.05 * exact, (blocks, exact)


@pytest.mark.parametrize("dims,blocks,blocks_y,tile_i,pencils", [((40, 30, 41), 3, 2, 16, 2), ((33, 64, 64), 2, 2, 16, 1), ((70, 17, 90), 8, 1, 32, 2),
                                                              ((16, 100, 9), 1, 4, 16, 1), ((50, 50, 50), 16, 16, 16, 2), ((20, 8, 8), 4, 4, 16, 2)])
def test_blocks_plan_covers_the_grid_and_cannot_deadlock(dims, blocks, blocks_y, tile_i, pencils):
    """The blocks ordering's host-side plan (s3d_sgs_blocks_plan: the layer tables and the dispatch order the tile kernel works from), walked on
    the CPU for both sweep directions: (1) every (row, plane) of the grid is STORED by exactly one tile row / layer, the not-stored overlap
    planes are the neighbouring block's; (2) a layer with a predecessor continues the layer before it in sweep order (the mailboxes are
    addressed by tile number +- ntx, +- ntx nty) and the first layer of a block has none: zero inflow; (3) the dispatch order is a
    topological order of the tile dependencies, so persistent warps that take tickets in order and wait for their predecessors cannot
    deadlock -- also played through with a few workers and random completion."""
    from struct3d_b200 import capi
    import random
    nx, ny, nz = dims
    for direction in (0, 1):
        (ntx, nty, ntz), kl, jl, order = capi.sgs_blocks_plan(nx, ny, nz, blocks, blocks_y, tile_i, pencils, direction)
        assert ntx == -(-nx // tile_i) and sorted(order.tolist()) == list(range(ntx * nty * ntz))
        for n, layers, T in ((nz, kl, 8), (ny, jl, 4 * pencils)):
            stored = np.zeros(n, dtype=int)
            for l, (first, cnt, flags, skip) in enumerate(layers.tolist()):
                assert 0 <= first and first + cnt <= n and 1 <= cnt <= T
                for q in range(first, first + cnt):
                    if q != skip:
                        stored[q] += 1
                pred, succ = flags & 1, (flags >> 1) & 1
                before = l - 1 if direction == 0 else l + 1          # the layer before this one in sweep order
                if pred:
                    assert 0 <= before < len(layers) and (layers[before][2] >> 1) & 1
                    lo, hi = (layers[before], layers[l]) if direction == 0 else (layers[l], layers[before])
                    assert lo[0] + lo[1] == hi[0], "layers of a block are contiguous"
                else:
                    # the first layer of a block in sweep order: it starts with the overlap plane unless the block touches the boundary
                    edge = first if direction == 0 else first + cnt - 1
                    assert skip == edge or edge in (0, n - 1)
            assert (stored == 1).all(), "every plane / row is stored exactly once"
        # dependencies of a tile: the tile before it along i (always, inside the grid), along j and k if the layer says so
        pos = np.empty(ntx * nty * ntz, dtype=int)
        seq = order if direction == 0 else order[::-1]
        pos[seq] = np.arange(len(seq))
        step = 1 if direction == 0 else -1
        preds = {}
        for t in range(ntx * nty * ntz):
            ci, cj, ck = t % ntx, (t // ntx) % nty, t // (ntx * nty)
            p = []
            if 0 <= ci - step < ntx:
                p.append(t - step)
            if jl[cj][2] & 1:
                p.append(t - step * ntx)
            if kl[ck][2] & 1:
                p.append(t - step * ntx * nty)
            preds[t] = p
            assert all(0 <= q < ntx * nty * ntz and pos[q] < pos[t] for q in p), "dispatch order is topological"
        rng = random.Random(7)
        for workers in (1, 3, 17):
            done, running, nxt = set(), {}, 0                      # running: worker -> tile
            for _ in range(20 * len(seq) + 100):
                for w in range(workers):
                    if w not in running and nxt < len(seq):
                        running[w] = int(seq[nxt]); nxt += 1
                ready = [w for w, t in running.items() if all(q in done for q in preds[t])]
                assert ready or not running, "deadlock"
                if not running:
                    break
                w = rng.choice(ready)
                done.add(running.pop(w))
            assert len(done) == len(seq)


def test_slab_grids_carry_slack_planes():
    """z-slab grids: every coefficient array has room for planes -1 .. nz (S3D_SGS_OVERLAP keeps the neighbours' boundary planes there);
    undecomposed grids are laid out as before"""
    from struct3d_b200 import capi
    g1 = capi.make_grid(40, 24, 30)
    assert g1.cplane * 30 <= g1.cstride < g1.cplane * 30 + 32
    for r in range(3):
        g = capi.make_grid(40, 24, 30, r, 3)
        assert g.cplane * (g.nz + 2) <= g.cstride < g.cplane * (g.nz + 2) + 32 and g.cstride % 32 == 0


def test_two_rank_slab_decomposition_gloo(tmp_path):
    """world_size-2 gloo run of the z-slab logic on CPU: the library's slab arithmetic + the halo protocol
    (top plane up, bottom plane down, Dirichlet zeros at the ends) reproduce the global SpMV bit for bit,
    slab partial dots allreduce to the global dot, and the SGS sweep pipelined through the slabs (forward: rank 0 then 1 with
    rank 0's top plane as rank 1's ghost; backward: rank 1 then 0) is the sequential sweep over the whole grid."""
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), S3D_PORT=str(port), S3D_ROOT=ROOT)
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    outs = [p.communicate(timeout=240)[0].decode() for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o

"""N > 1 host logic on CPU (world_size 2 and 3, gloo): the z-slab plan of b200md_pppm_decomp is replayed with numpy
arrays standing in for device buffers — density halo sum, slab FFT with the two all-to-all transposes and the field
halo fill, exchanged over torch.distributed exactly as csrc/pppm.cu exchanges them over NCCL — and must reproduce
the single-process result.  Also: the per-rank bench blocks tile the global box."""
import ctypes as C
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent('''
    import ctypes as C, os, sys
    import numpy as np
    import torch
    import torch.distributed as dist
    sys.path.insert(0, %(root)r)
    import __graft_entry__ as graft
    pkg = graft.load_package()
    lib = pkg.load()
    rank, P = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    nx, ny, nz, order, skin, Lz = 12, 10, %(nz)d, 5, 0.6, 40.0
    ia = lambda: (C.c_int * P)()
    pzlo, pzhi, zoff, nbz, ylo, yhi = ia(), ia(), ia(), ia(), ia(), ia()
    rc = lib.b200md_pppm_decomp(C.c_int(P), C.c_int(nz), C.c_int(ny), C.c_int(order), C.c_double(skin), C.c_double(Lz),
                                pzlo, pzhi, zoff, nbz, ylo, yhi)
    assert rc == 0
    me, lower, upper = rank, (rank - 1) %% P, (rank + 1) %% P
    # plan invariants
    assert pzlo[0] == 0 and pzhi[P - 1] == nz and all(pzhi[r] == pzlo[r + 1] for r in range(P - 1))
    assert ylo[0] == 0 and yhi[P - 1] == ny and all(yhi[r] == ylo[r + 1] for r in range(P - 1))
    assert all(zoff[r] <= pzlo[r] and zoff[r] + nbz[r] >= pzhi[r] for r in range(P))
    lo_w = lambda r: pzlo[r] - zoff[r]
    hi_w = lambda r: zoff[r] + nbz[r] - pzhi[r]
    # a global "density" whose brick-local contributions are known: every rank deposits g(z) * w_r on the planes of its
    # brick (periodic), the owners must end up with the sum over the ranks whose bricks cover the plane
    rng = np.random.default_rng(7)
    base = rng.normal(size=(nz, ny, nx))
    brick_planes = (np.arange(nbz[me]) + zoff[me]) %% nz
    local = base[brick_planes] * (1.0 + me)                     # this rank's brick [nbz][ny][nx]
    nzo = pzhi[me] - pzlo[me]
    own = local[lo_w(me):lo_w(me) + nzo].copy()

    def exchange(s_lo, s_hi, n_from_hi, n_from_lo):
        """b2_comm_exchange: send to lower / upper, receive from upper / lower (message order as in comm.cu)"""
        r_hi = torch.zeros(n_from_hi, dtype=torch.float64)
        r_lo = torch.zeros(n_from_lo, dtype=torch.float64)
        ops = []
        if s_lo.numel(): ops.append(dist.P2POp(dist.isend, s_lo, lower))
        if s_hi.numel(): ops.append(dist.P2POp(dist.isend, s_hi, upper))
        if n_from_hi: ops.append(dist.P2POp(dist.irecv, r_hi, upper))
        if n_from_lo: ops.append(dist.P2POp(dist.irecv, r_lo, lower))
        for w in dist.batch_isend_irecv(ops): w.wait()
        return r_hi.numpy(), r_lo.numpy()

    plane = ny * nx
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a).ravel().copy())
    # ---- density halo sum ----
    r_hi, r_lo = exchange(t(local[:lo_w(me)]), t(local[lo_w(me) + nzo:]), lo_w(upper) * plane, hi_w(lower) * plane)
    if hi_w(lower): own[:hi_w(lower)] += r_lo.reshape(-1, ny, nx)
    if lo_w(upper): own[nzo - lo_w(upper):] += r_hi.reshape(-1, ny, nx)
    expect = np.zeros((nz, ny, nx))
    for r in range(P):
        for l in range(nbz[r]):
            expect[(zoff[r] + l) %% nz] += base[(zoff[r] + l) %% nz] * (1.0 + r)
    assert np.allclose(own, expect[pzlo[me]:pzhi[me]], rtol=0, atol=1e-12), "density halo sum"
    # ---- slab FFT: x, y local; transpose; z; compare with the global transform ----
    w1 = np.fft.fft(np.fft.fft(own.astype(complex), axis=2), axis=1)            # [nzo][ny][nx]
    send = [torch.from_numpy(np.ascontiguousarray(w1[:, ylo[q]:yhi[q], :]).view(np.float64).ravel().copy()) for q in range(P)]
    recv = [torch.zeros((pzhi[q] - pzlo[q]) * (yhi[me] - ylo[me]) * nx * 2, dtype=torch.float64) for q in range(P)]
    dist.all_to_all(recv, send) if dist.get_backend() != "gloo" else [
        w.wait() for w in dist.batch_isend_irecv([dist.P2POp(dist.isend, send[q], q) for q in range(P) if q != me] +
                                                 [dist.P2POp(dist.irecv, recv[q], q) for q in range(P) if q != me])]
    recv[me] = send[me]
    nyl = yhi[me] - ylo[me]
    wT = np.concatenate([r.numpy().view(complex).reshape(pzhi[q] - pzlo[q], nyl, nx) for q, r in enumerate(recv)], axis=0)
    wT = np.fft.fft(wT, axis=0)                                                  # [nz][nyl][nx]
    ref = np.fft.fftn(expect)
    assert np.allclose(wT, ref[:, ylo[me]:yhi[me], :], rtol=1e-12, atol=1e-9), "forward slab FFT"
    # ---- back: inverse z, transpose (contiguous z chunks out, row blocks in), inverse y, x ----
    wT = np.fft.ifft(wT, axis=0)
    send = [torch.from_numpy(np.ascontiguousarray(wT[pzlo[q]:pzhi[q]]).view(np.float64).ravel().copy()) for q in range(P)]
    recv = [torch.zeros(nzo * (yhi[q] - ylo[q]) * nx * 2, dtype=torch.float64) for q in range(P)]
    [w.wait() for w in dist.batch_isend_irecv([dist.P2POp(dist.isend, send[q], q) for q in range(P) if q != me] +
                                              [dist.P2POp(dist.irecv, recv[q], q) for q in range(P) if q != me])]
    recv[me] = send[me]
    w2 = np.concatenate([r.numpy().view(complex).reshape(nzo, yhi[q] - ylo[q], nx) for q, r in enumerate(recv)], axis=1)
    back = np.fft.ifft(np.fft.ifft(w2, axis=1), axis=2).real
    assert np.allclose(back, expect[pzlo[me]:pzhi[me]], rtol=0, atol=1e-10), "inverse slab FFT"
    # ---- field halo fill: the brick is owned planes + the neighbours' boundary planes ----
    r_hi, r_lo = exchange(t(back[:hi_w(lower)]), t(back[nzo - lo_w(upper):] if lo_w(upper) else back[:0]),
                          hi_w(me) * plane, lo_w(me) * plane)
    brick = np.concatenate([r_lo.reshape(-1, ny, nx), back, r_hi.reshape(-1, ny, nx)], axis=0)
    assert brick.shape[0] == nbz[me]
    assert np.allclose(brick, expect[brick_planes], rtol=0, atol=1e-10), "field halo fill"
    dist.barrier()
    if rank == 0: print("DECOMP OK", P)
''')


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,nz", [(2, 24), (3, 36), (2, 27)])
def test_slab_plan_replayed_over_gloo(tmp_path, world, nz):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % dict(root=ROOT, nz=nz))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), str(script)]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "DECOMP OK %d" % world in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_decomp_rejects_thin_slabs(pkg):
    lib = pkg.load()
    P = 8
    arr = [(C.c_int * P)() for _ in range(6)]
    # 16 planes over 8 ranks: 2 owned planes per rank cannot hold an order-5 stencil halo
    assert lib.b200md_pppm_decomp(C.c_int(P), C.c_int(16), C.c_int(16), C.c_int(5), C.c_double(0.3), C.c_double(30.0), *arr) != 0
    assert lib.b200md_pppm_decomp(C.c_int(P), C.c_int(270 * 8), C.c_int(250), C.c_int(5), C.c_double(0.3), C.c_double(3362.0), *arr) == 0
    assert arr[0][0] == 0 and arr[1][P - 1] == 270 * 8


def test_bench_rank_blocks_tile_the_global_system(W):
    """bench.py at N > 1: every rank generates the z blocks overlapping its slab and keeps the atoms inside it; together
    they are the replicated crystal — also when the blocks do not align with the slabs (cubic geometry: 5 cells over
    3 ranks)"""
    import importlib
    sys.path.insert(0, ROOT)
    bench = importlib.import_module("bench")
    make = lambda r3: W.aC_system(r3, jitter=0.0)
    for P, reps in ((3, (2, 2, 6)), (3, (2, 2, 5)), (2, (1, 1, 3))):
        blocks = [bench.rank_system(make, reps, r, P) for r in range(P)]
        whole = W.aC_system(reps, jitter=0.0)
        x = np.concatenate([b["x"] for b in blocks])
        assert np.allclose(blocks[0]["boxhi"], whole["boxhi"]) and len(x) == len(whole["x"])
        key = lambda a: np.lexsort(np.round(a, 6).T[::-1])
        assert np.allclose(x[key(x)], whole["x"][key(whole["x"])], atol=1e-9)
        lz = (whole["boxhi"][2] - whole["boxlo"][2]) / P
        for r, b in enumerate(blocks):
            assert (b["x"][:, 2] >= r * lz - 1e-9).all() and (b["x"][:, 2] < (r + 1) * lz + 1e-9).all()
            assert len(b["x"]) == len(b["type"]) == len(b["q"]) == len(b["v"])


def test_bench_cube_geometry_keeps_the_per_gpu_atom_count():
    import importlib
    sys.path.insert(0, ROOT)
    bench = importlib.import_module("bench")
    for world, r in ((1, 15), (2, 19), (4, 24), (8, 30)):      # data.aC x 30^3 on 8 GPUs
        assert bench.split_reps(15, world, "cube", "weak") == (r, r, r)
        assert abs(r ** 3 / world / 15 ** 3 - 1.0) < 0.03
    assert bench.split_reps(15, 8, "slab", "weak") == (15, 15, 120)
    assert bench.split_reps(24, 4, "slab", "strong") == (24, 24, 24)


def _memcpy2d(dst, doff, dpitch, src, soff, spitch, width, height):
    """cudaMemcpy2DAsync on flat arrays (element units): `height` rows of `width` elements"""
    for r in range(height):
        dst[doff + r * dpitch: doff + r * dpitch + width] = src[soff + r * spitch: soff + r * spitch + width]


@pytest.mark.parametrize("P,nx,ny,nz,npack", [(2, 6, 10, 24, 2), (3, 5, 10, 37, 2), (4, 4, 9, 48, 1), (8, 3, 13, 96, 2),
                                                (8, 7, 13, 96, 3), (4, 4, 10, 50, 3)])   # npack 3: the half-spectrum path (nx = sx)
def test_peer_copy_plan_tiles_the_transposes(pkg, P, nx, ny, nz, npack):
    """the copy-engine transposes of csrc/pppm.cu (poisson_multi): the strided copies every rank issues into its peers'
    pencil / plane blocks — destination offset and pitch, source offset and pitch, width, height exactly as in the
    cudaMemcpy2DAsync calls — cover each destination block once and put every element where the next pass reads it
    (uneven row / plane splits included)"""
    lib = pkg.load()
    arr = [(C.c_int * P)() for _ in range(6)]
    assert lib.b200md_pppm_decomp(C.c_int(P), C.c_int(nz), C.c_int(ny), C.c_int(5), C.c_double(0.3), C.c_double(4.0 * nz),
                                  *arr) == 0
    pzlo, pzhi, _, _, ylo, yhi = [list(a) for a in arr]
    plane = nx * ny
    rng = np.random.default_rng(3)
    G = rng.normal(size=(nz, ny, nx))                       # the global array after the x / y passes
    maxrows = max(yhi[q] - ylo[q] for q in range(P))
    maxplanes = max(pzhi[q] - pzlo[q] for q in range(P))
    symT = [np.full(nx * maxrows * nz, np.nan) for _ in range(P)]
    # forward: rank me holds its planes [nzo][ny][nx]; one copy per destination q
    for me in range(P):
        nzo = pzhi[me] - pzlo[me]
        work1 = np.ascontiguousarray(G[pzlo[me]:pzhi[me]]).ravel()
        for k in range(P):
            q = (me + k) % P
            nylq = yhi[q] - ylo[q]
            if nylq == 0 or nzo == 0:
                continue
            _memcpy2d(symT[q], pzlo[me] * nylq * nx, nylq * nx, work1, ylo[q] * nx, plane, nylq * nx, nzo)
    for q in range(P):
        nyl = yhi[q] - ylo[q]
        got = symT[q][:nz * nyl * nx].reshape(nz, nyl, nx)
        assert np.array_equal(got, G[:, ylo[q]:yhi[q], :]), ("forward", q)
    # backward: rank me holds its pencils [npack][nz][nyl][nx]; one copy per destination and packed field
    H = [rng.normal(size=(nz, ny, nx)) for _ in range(npack)]
    symW = [np.full(npack * plane * maxplanes, np.nan) for _ in range(P)]
    for me in range(P):
        nyl = yhi[me] - ylo[me]
        nT = nx * nyl * nz
        workT2 = np.concatenate([np.ascontiguousarray(H[c][:, ylo[me]:yhi[me], :]).ravel() for c in range(npack)])
        for k in range(P):
            q = (me + k) % P
            nzq = pzhi[q] - pzlo[q]
            if nzq == 0 or nyl == 0:
                continue
            for comp in range(npack):
                _memcpy2d(symW[q], (comp * nzq * ny + ylo[me]) * nx, plane, workT2, comp * nT + pzlo[q] * nyl * nx,
                          nyl * nx, nyl * nx, nzq)
    for q in range(P):
        nzq = pzhi[q] - pzlo[q]
        got = symW[q][:npack * nzq * plane].reshape(npack, nzq, ny, nx)
        for c in range(npack):
            assert np.array_equal(got[c], H[c][pzlo[q]:pzhi[q]]), ("backward", q, c)

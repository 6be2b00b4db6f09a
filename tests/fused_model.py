"""CPU model of the fused kernel (intel-qs_b200/csrc/kernels_fused.cu: k_fused), test infrastructure.

It executes the RAW descriptors iqsb_fused hands to the kernel (iqsb_plan_fused_dump): tile positions,
group headers (swizzled slot tables of the threads, load basis p, write-back basis w / c0), gates with
their class, register bit, control kind and matrix -- with the kernel's own rules:

  * a tile = the 2^nS amplitudes whose indices differ in the positions pos[0..nS);
  * thread t of a group owns the 8 amplitudes at slots px ^ span_p(r), r = 0..7, px from the tables;
  * main gates act on register pairs (r0, r0 | 1 << tbit), pair k enabled by bit k of `en`; a control on
    a thread bit / a bit of the tile's base index switches the whole gate off for that thread / tile;
  * arithmetic per matrix class exactly as the kernel writes it (separately rounded products);
  * write-back to slots px ^ c0 ^ [conditional offsets] ^ span_w(r).

So the host side of fusion -- runs, groups, exact-commutation reordering, folding of X / CNOT into the
write-back addresses, class assignment -- is checked against the oracle on the CPU, without a GPU.
"""
import numpy as np

N_GROUPS, N_GATES = 24, 48
GROUP_DT = np.dtype([("lo", "<u2", 32), ("hi", "<u2", 16), ("p", "<u2", 4), ("gate_first", "<u2"), ("gate_count", "<u2"),
                     ("log2_threads", "<u2"), ("nmain", "<u2"), ("ncond", "<u2"), ("w", "<u2", 4), ("c0", "<u2"), ("pad", "<u2", 2)])
GATE_DT = np.dtype([("m", "<f8", 8), ("cls", "u1"), ("tbit", "u1"), ("ckind", "u1"), ("c", "u1"), ("en", "u1"), ("last", "u1"),
                    ("trail", "u1"), ("pad8", "u1"), ("origin", "<u4"), ("pu", "<u2"), ("pad16", "<u2")])
assert GROUP_DT.itemsize == 128 and GATE_DT.itemsize == 80
BATCH_BYTES = 16 + N_GROUPS * GROUP_DT.itemsize + N_GATES * GATE_DT.itemsize
GENERAL, REAL, DIAG, DIAG1, ANTI, XEXACT, RX, SQRTX, SQRTY = range(9)


def swz(i):
    """the kernel's shared-memory swizzle; an involution, linear over GF(2)"""
    i = np.asarray(i, dtype=np.int64)
    return i ^ (((i >> 3) ^ (i >> 6) ^ (i >> 9)) & 7)


def parse(blob):
    """-> [ {nS, pos, batches: [ {groups: structured array, gates: structured array} ]} ]"""
    off = 0
    nruns = int(np.frombuffer(blob, "<i4", 1, off)[0]); off += 4
    runs = []
    for _ in range(nruns):
        nS = int(np.frombuffer(blob, "<i4", 1, off)[0]); off += 4
        pos = np.frombuffer(blob, "u1", 12, off).astype(int)[:nS]; off += 12
        nb = int(np.frombuffer(blob, "<i4", 1, off)[0]); off += 4
        batches = []
        for _ in range(nb):
            ngroups, ngates = (int(x) for x in np.frombuffer(blob, "<i4", 2, off))
            groups = np.frombuffer(blob, GROUP_DT, N_GROUPS, off + 16)[:ngroups]
            gates = np.frombuffer(blob, GATE_DT, N_GATES, off + 16 + N_GROUPS * GROUP_DT.itemsize)[:ngates]
            batches.append(dict(groups=groups, gates=gates))
            off += BATCH_BYTES
        runs.append(dict(nS=nS, pos=pos, batches=batches))
    assert off == len(blob)
    return runs


def _cmul(mr, mi, x):
    """(mr + i mi) * x with separately rounded products, the kernel's cmul"""
    return (mr * x.real - mi * x.imag) + 1j * (mr * x.imag + mi * x.real)


def _apply_class(cls, m, x, y):
    """out0, out1 of one register pair; formulas as in apply_on_bit (exact arithmetic mode)"""
    m00, m01, m10, m11 = (complex(m[0], m[1]), complex(m[2], m[3]), complex(m[4], m[5]), complex(m[6], m[7]))
    if cls == XEXACT:
        return y, x
    if cls == DIAG1:
        return x, _cmul(m11.real, m11.imag, y)
    if cls == DIAG:
        return _cmul(m00.real, m00.imag, x), _cmul(m11.real, m11.imag, y)
    if cls == ANTI:
        return _cmul(m01.real, m01.imag, y), _cmul(m10.real, m10.imag, x)
    if cls == REAL:
        o0 = (m00.real * x.real + m01.real * y.real) + 1j * (m00.real * x.imag + m01.real * y.imag)
        o1 = (m10.real * x.real + m11.real * y.real) + 1j * (m10.real * x.imag + m11.real * y.imag)
        return o0, o1
    if cls == RX:
        o0 = (m00.real * x.real - m01.imag * y.imag) + 1j * (m00.real * x.imag + m01.imag * y.real)
        o1 = (-(m10.imag * x.imag) + m11.real * y.real) + 1j * (m10.imag * x.real + m11.real * y.imag)
        return o0, o1
    if cls in (SQRTX, SQRTY):
        A, B, Cp, D = x.real - x.imag, x.real + x.imag, y.real + y.imag, y.real - y.imag
        if cls == SQRTX:
            return 0.5 * (A + Cp) + 1j * (0.5 * (B - D)), 0.5 * (B + D) + 1j * (0.5 * (Cp - A))
        return 0.5 * (A - D) + 1j * (0.5 * (B - Cp)), 0.5 * (A + D) + 1j * (0.5 * (B + Cp))
    return (_cmul(m00.real, m00.imag, x) + _cmul(m01.real, m01.imag, y), _cmul(m10.real, m10.imag, x) + _cmul(m11.real, m11.imag, y))


def _span(basis, r):
    out = 0
    for j in range(3):
        if (r >> j) & 1:
            out ^= int(basis[j])
    return out


def run(state, blob, log2_local):
    """Execute the schedule on a copy of `state` (2^log2_local complex128); returns the new state."""
    psi = np.array(state, dtype=np.complex128, copy=True)
    n = log2_local
    for rn in parse(blob):
        nS, pos = rn["nS"], rn["pos"]
        assert pos[0] == 0 and list(pos) == sorted(set(pos))
        slots = np.arange(1 << nS, dtype=np.int64)
        off = np.zeros_like(slots)
        for k in range(nS):
            off |= ((slots >> k) & 1) << int(pos[k])
        others = [b for b in range(n) if b not in set(int(p) for p in pos)]
        outer = np.arange(1 << len(others), dtype=np.int64)
        base = np.zeros_like(outer)
        for k, b in enumerate(others):
            base |= ((outer >> k) & 1) << b
        idx = base[:, None] | off[None, :]
        tile = psi[idx]  # (tiles, 2^nS)
        for batch in rn["batches"]:
            groups, gates = batch["groups"], batch["gates"]
            for G in groups:
                T = 1 << int(G["log2_threads"])
                assert int(G["log2_threads"]) == nS - 3
                t = np.arange(T, dtype=np.int64)
                px = swz(G["lo"][t & 31].astype(np.int64) ^ G["hi"][t >> 5].astype(np.int64))  # logical slot of register 0
                P = [int(swz(int(v))) for v in G["p"][:3]]
                W = [int(swz(int(v))) for v in G["w"][:3]]
                load = np.stack([px ^ _span(P, r) for r in range(8)], axis=1)  # (T, 8)
                assert len(set(load.ravel().tolist())) == T * 8 == 1 << nS  # the threads partition the tile
                a = tile[:, load]  # (tiles, T, 8)
                first, nmain, ncond = int(G["gate_first"]), int(G["nmain"]), int(G["ncond"])
                for gj in range(first, first + nmain):
                    g = gates[gj]
                    assert g["trail"] == 0 and (g["last"] == 1) == (gj == first + nmain - 1)
                    tb, ck, c = int(g["tbit"]), int(g["ckind"]), int(g["c"])
                    # the arithmetic classes are straight-line code over all four pairs: only X honours `en`
                    assert int(g["cls"]) == XEXACT or (int(g["en"]) == 0xF and ck != 1)
                    active = np.ones((tile.shape[0], T), dtype=bool)
                    if ck == 2:
                        active &= (((t >> c) & 1) == 1)[None, :]
                    elif ck == 3:
                        active &= (((base >> c) & 1) == 1)[:, None]
                    for k in range(4):
                        if not (int(g["en"]) >> k) & 1:
                            continue
                        r0 = ((k >> tb) << (tb + 1)) | (k & ((1 << tb) - 1))
                        r1 = r0 | (1 << tb)
                        x, y = a[:, :, r0].copy(), a[:, :, r1].copy()
                        o0, o1 = _apply_class(int(g["cls"]), g["m"], x, y)
                        a[:, :, r0] = np.where(active, o0, x)
                        a[:, :, r1] = np.where(active, o1, y)
                pxs = np.broadcast_to(px ^ int(swz(int(G["c0"]))), (tile.shape[0], T)).copy()
                for gj in range(first + nmain, first + nmain + ncond):
                    g = gates[gj]
                    assert g["trail"] == 1 and g["cls"] == XEXACT and g["ckind"] in (2, 3)
                    c = int(g["c"])
                    on = (((t >> c) & 1) == 1)[None, :] if g["ckind"] == 2 else (((base >> c) & 1) == 1)[:, None]
                    pxs ^= np.where(np.broadcast_to(on, pxs.shape), int(swz(int(g["pu"]))), 0)
                for gj in range(first + nmain + ncond, first + int(G["gate_count"])):
                    assert gates[gj]["trail"] == 2 and gates[gj]["cls"] == XEXACT  # absorbed into (w, c0)
                store = np.stack([pxs ^ _span(W, r) for r in range(8)], axis=2)  # (tiles, T, 8)
                flat = np.sort(store.reshape(tile.shape[0], -1), axis=1)
                assert np.array_equal(flat, np.broadcast_to(slots, flat.shape))  # the write-back is a permutation of the tile
                rows = np.arange(tile.shape[0])[:, None, None]
                new_tile = np.empty_like(tile)
                new_tile[rows, store] = a
                tile = new_tile
        psi[idx] = tile
    return psi

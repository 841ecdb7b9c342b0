"""Static properties of the TMA-staged GEMM's shared-memory addressing (afesp_b200/csrc/gemm_tma.cu), checked by
enumeration on the CPU: the k-index sets {0,3,12,15} {1,2,13,14} {4,7,8,11} {5,6,9,10} make every DMMA fragment load
bank-conflict free in BOTH 128-byte-swizzled tile layouts TMA produces (MN-major boxes [mn/16][k][16] and K-major rows
[mn][16 k]), and the natural choice {0..3} {4..7} ... does not.  Mirrors tile_off<KMAJOR>() and kslot()."""
import itertools


def tile_off(kmajor, mn, k):
    if kmajor:
        return mn * 128 + ((((k >> 1) ^ (mn & 7)) << 4) | ((k & 1) << 3))
    return (mn >> 4) * 2048 + k * 128 + (((((mn & 15) >> 1) ^ (k & 7)) << 4) | ((mn & 1) << 3))


def kslot(s, t):
    return ((1 + (t & 1)) if (s & 1) else 3 * (t & 1)) + ((((s >> 1) ^ 1) * 4 + 8) if (t >> 1) else (s >> 1) * 4)


def _half_warp_conflicts(kmajor, kfun):
    """max number of lanes of one half-warp (16 lanes x 8 B = one 128-byte shared-memory wavefront) that fall on the
    same 8-byte bank pair, over all k-sets, fragment rows and warp positions."""
    worst = 1
    for s, i, w0 in itertools.product(range(4), range(4), (0, 32)):
        for half in (0, 1):
            banks = {}
            for lane in range(16 * half, 16 * half + 16):
                gid, tig = lane >> 2, lane & 3
                off = tile_off(kmajor, w0 + 8 * i + gid, kfun(s, tig))
                assert off % 8 == 0 and 0 <= off < 64 * 16 * 8
                b = (off // 8) % 16
                banks[b] = banks.get(b, 0) + 1
            worst = max(worst, max(banks.values()))
    return worst


def test_kslot_sets_partition_the_k_tile():
    ks = sorted(kslot(s, t) for s in range(4) for t in range(4))
    assert ks == list(range(16))
    assert [sorted(kslot(s, t) for t in range(4)) for s in range(4)] == [[0, 3, 12, 15], [1, 2, 13, 14], [4, 7, 8, 11],
                                                                         [5, 6, 9, 10]]


def test_tile_offsets_are_a_bijection_onto_the_8kb_tile():
    for kmajor in (False, True):
        offs = sorted(tile_off(kmajor, mn, k) for mn in range(64) for k in range(16))
        assert offs == list(range(0, 64 * 16 * 8, 8))


def test_fragment_loads_are_bank_conflict_free_in_both_layouts():
    assert _half_warp_conflicts(False, kslot) == 1
    assert _half_warp_conflicts(True, kslot) == 1


def test_the_natural_k_grouping_would_conflict():
    natural = lambda s, t: 4 * s + t
    assert max(_half_warp_conflicts(False, natural), _half_warp_conflicts(True, natural)) >= 2


def test_fused_triples_epilogue_box_swizzles_are_conflict_free():
    """afesp_b200/csrc/triples.cu `swz`: each staged 8x8x8 box is written at the thread's natural position and read back
    at the position permuted by w; in its own XOR-swizzled layout both accesses of every half-warp (tx = 0..7, two
    consecutive ty, fixed tz) touch 16 distinct 8-byte banks, and the layout is a bijection onto the 512-word box."""
    perms = [(0, 1, 2), (1, 0, 2), (2, 1, 0), (0, 2, 1), (1, 2, 0), (2, 0, 1)]   # c_perm of triples.cu
    swz = {1: lambda X, Y, Z: 64 * Z + 8 * Y + (X ^ Y),
           2: lambda X, Y, Z: 64 * Z + 8 * Y + (X ^ Z),
           3: lambda X, Y, Z: 64 * Z + 8 * (Y ^ (Z & 1)) + X,
           4: lambda X, Y, Z: 64 * Z + 8 * (Y ^ (X & 1)) + (X ^ Z),
           5: lambda X, Y, Z: 64 * Z + 8 * (Y ^ (Z & 1)) + (X ^ Y)}
    for w, f in swz.items():
        assert sorted(f(x, y, z) for x in range(8) for y in range(8) for z in range(8)) == list(range(512))
        p = perms[w]
        for tz in range(8):
            for t0 in range(0, 8, 2):
                half = [(tx, ty, tz) for ty in (t0, t0 + 1) for tx in range(8)]
                assert len({f(*l) % 16 for l in half}) == 16, ("write", w)
                assert len({f(l[p[0]], l[p[1]], l[p[2]]) % 16 for l in half}) == 16, ("read", w)
    # the padded 9x9x8 box of the previous version conflicted 2-way on every access, natural ones included
    old = lambda X, Y, Z: (Z * 9 + Y) * 9 + X
    assert len({old(tx, ty, 0) % 16 for ty in (0, 1) for tx in range(8)}) == 15

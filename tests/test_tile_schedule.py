"""CPU restatement of the GEMM's tile schedule (m3p_b200/csrc/gemm.cu: gemm_impl + decode_unit): with tail balancing
the last partial wave of 256 x 256 tiles is issued as half-width units.  Whatever the shape, every output element must
be produced by exactly one unit, and the units a CTA pair walks (static stride) must cost at most what the uniform
schedule cost."""
import pytest

BN, TILE_M, WORKERS = 256, 256, 74  # CTA pairs on a 148-SM B200


def schedule(m, n, tail_balance=True):
    """-> list of units (m_tile, n0, bn) in launch order, exactly as decode_unit() enumerates them (split_k == 1)."""
    n_m, n_n = -(-m // TILE_M), -(-n // BN)
    m_fastest = n_m < n_n
    units = n_m * n_n
    wide = units
    if tail_balance and not m_fastest and n % BN == 0:
        rem = units % WORKERS
        if units > WORKERS and 0 < rem and 2 * rem <= WORKERS:
            wide, units = units - rem, units + rem
    out = []
    for u in range(units):
        if u >= wide:
            g = 2 * wide + (u - wide)
            per_row = 2 * n_n
            out.append((g // per_row, (g % per_row) * (BN // 2), BN // 2))
        else:
            mt, nt = (u % n_m, u // n_m) if m_fastest else (u // n_n, u % n_n)
            out.append((mt, nt * BN, BN))
    return out


SHAPES = [(14592, 768), (14592, 3072), (14592, 2304), (7296, 768), (6600, 768), (1792, 3072), (14592, 1024), (14592, 4096),
          (300, 392), (1024, 250008), (64, 768), (19200, 768)]


@pytest.mark.parametrize("m,n", SHAPES)
def test_every_output_tile_is_covered_exactly_once(m, n):
    seen = {}
    for mt, n0, bn in schedule(m, n):
        assert n0 % (BN // 2) == 0 and bn in (BN, BN // 2)
        for half in range(n0 // (BN // 2), (n0 + bn) // (BN // 2)):
            key = (mt, half)
            assert key not in seen, key
            seen[key] = True
    n_m = -(-m // TILE_M)
    halves = 2 * (-(-n // BN))
    assert len(seen) == n_m * halves and all(0 <= mt < n_m and 0 <= h < halves for mt, h in seen)


@pytest.mark.parametrize("m,n", SHAPES)
def test_balanced_schedule_never_costs_a_cta_pair_more(m, n):
    def makespan(units):
        cost = [0.0] * WORKERS
        for u, (_, _, bn) in enumerate(units):
            cost[u % WORKERS] += bn / BN
        return max(cost)

    uni, bal = makespan(schedule(m, n, False)), makespan(schedule(m, n, True))
    assert bal <= uni
    if (m, n) == (14592, 768):  # the encoder's N = 768 GEMMs: 171 tiles = 3 waves -> 2.5 tile times
        assert (uni, bal) == (3.0, 2.5)

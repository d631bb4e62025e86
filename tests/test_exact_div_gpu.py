"""The rasteriser's shared-reciprocal division (csrc/exact.cuh: Recip / xdiv_by) must give the bits of div.rn.f32 --
which is what the reference's `/` compiles to (rasteriser.cpp:538-541, :622-624, :648-649, :557) -- for every operand
pair, inside and outside its fast-path guard."""
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_shared_reciprocal_division_matches_div_rn(pkg, seed):
    ctx = pkg.Context(8, 8)
    bad, first = ctx.selftest_division(1 << 31, seed)  # 2.1e9 pairs per seed, ~1 s on a B200
    ctx.close()
    assert bad == 0, f"{bad} mismatches; first: a={first[0]!r} b={first[1]!r} div.rn={first[2]!r} shared={first[3]!r}"

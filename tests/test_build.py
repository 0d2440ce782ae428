"""CPU tier: properties of the compiled sm_100a code that the exactness argument relies on
(DESIGN.md "Exactness"): no fused multiply-add where the oracle rounds twice."""
import os
import re
import subprocess

import pytest

from conftest import ROOT

LIB = os.path.join(ROOT, "gr-bluetooth_b200", "libbtb200.so")


@pytest.fixture(scope="module")
def sass():
    if not os.path.exists(LIB):
        pytest.skip("libbtb200.so not built")
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    funcs = {}
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
        elif cur and "/*" in line:
            funcs[cur].append(line)
    return funcs


def body(funcs, key):
    names = [n for n in funcs if key in n]
    assert names, key
    return [l for n in names for l in funcs[n]]


def test_built_for_sm100a():
    out = subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_exact_fir_kernels_have_no_fused_multiply_add(sass):
    for key in ("k_fir_tiled", "k_chan_fir_v1", "k_noise_fir_v1"):
        lines = body(sass, key)
        assert any(" FMUL " in l for l in lines) and any(" FADD " in l for l in lines)
        assert not any(re.search(r"\bFFMA\b", l) for l in lines), key


def test_packed_fir_keeps_separate_roundings(sass):
    """ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2; the kernel avoids that pattern.
    Every FFMA2 left must be the 'x - p' form with the -1 immediate (exact), and the
    instruction mix per complex MAC pair is 4 FMUL2 : 2 FFMA2 : 2 FADD2."""
    lines = body(sass, "k_fir_packed")
    ffma2 = [l for l in lines if " FFMA2 " in l]
    assert ffma2 and all(", -1, " in l for l in ffma2)
    nmul = sum(" FMUL2 " in l for l in lines)
    nadd = sum(" FADD2 " in l for l in lines)
    assert nmul == 2 * len(ffma2) and nadd == len(ffma2)
    assert not any(re.search(r"\bFFMA\b", l) for l in lines)
    assert any("LDS.128" in l for l in lines)
    assert any("UBLKCP" in l for l in lines)          # tap banks arrive by TMA bulk copy (cp.async.bulk + mbarrier)


def test_delay_line_fir_keeps_separate_roundings(sass):
    """Same rule for the delay-line noise FIR (rx_firdl.cu): FFMA2 only as 'x - p' with the -1
    immediate, 4 FMUL2 : 2 FFMA2 : 2 FADD2; inputs fetched with cp.async (LDGSTS), tap banks with TMA bulk copies."""
    lines = body(sass, "k_fir_dl")
    ffma2 = [l for l in lines if " FFMA2 " in l]
    assert ffma2 and all(", -1, " in l for l in ffma2)
    nmul = sum(" FMUL2 " in l for l in lines)
    nadd = sum(" FADD2 " in l for l in lines)
    assert nmul == 2 * len(ffma2) and nadd == len(ffma2)
    assert not any(re.search(r"\bFFMA\b", l) for l in lines)
    assert any("LDGSTS" in l for l in lines)                      # cp.async: input ring
    assert any("UBLKCP" in l for l in lines)                      # cp.async.bulk (TMA): tap banks
    assert any("SYNCS.ARRIVE.TRANS64" in l for l in lines)        # mbarrier expect_tx


def test_no_fmad_flag_in_makefile():
    mk = open(os.path.join(ROOT, "gr-bluetooth_b200", "Makefile")).read()
    assert "--fmad=false" in mk and "-ffp-contract=off" in mk


def test_throughput_mode_kernels_are_tma_fed_fma_kernels(sass):
    """The polyphase channelizer and the noise estimator (rx_pfb.cu, rx_nest.cu): IQ tiles, sample ring and tables
    arrive by TMA bulk copies on mbarriers (SASS UBLKCP / SYNCS), the arithmetic is fused multiply-add -- packed FFMA2 in the
    estimator's tap loop -- and no tensor-core instruction appears anywhere (north_star: short FIRs, not a contraction)."""
    pfb = body(sass, "k_pfbILi4ELi7")
    assert any("UBLKCP" in l for l in pfb) and any("SYNCS.ARRIVE.TRANS64" in l for l in pfb)
    assert sum(bool(re.search(r"\bFFMA\b", l)) for l in pfb) > 300
    nest = body(sass, "k_nestILi4ELi100")
    assert any("UBLKCP" in l for l in nest) and any("SYNCS.PHASECHK" in l for l in nest)
    assert sum(" FFMA2 " in l for l in nest) >= 256
    for name, lines in sass.items():
        assert not any(re.search(r"\b(HMMA|UTC\w*MMA|LDTM|STTM)\b", l) for l in lines), name


def test_clock_recovery_is_not_contracted_where_it_rides_in_the_estimator(sass):
    """rx_nest.cu also carries the resume of the Mueller & Mueller loop (rx_mm.cuh): that code must keep one rounding per
    operation (the 8-tap dot product is FMUL + FADD), so the file is built without contraction and its FMAs are explicit."""
    mk = open(os.path.join(ROOT, "gr-bluetooth_b200", "Makefile")).read()
    rule = mk[mk.index("build/rx_nest.o:"):]
    assert "$(EXACT)" in rule.split("build/plan.o")[0] and "$(FAST)" not in rule.split("build/plan.o")[0].split("\n", 3)[2]
    mm = body(sass, "k_mm_stateless_v2")
    assert any(" FMUL " in l for l in mm) and any(" FADD " in l for l in mm)
    # the only fused multiply-adds are the exact ones of the table index (rx_mm.cuh: 128 mu + 1.5 2^23 with the floor
    # folded in: both products are exact, one rounding, same value as the reference's rint(128 mu))
    for l in mm:
        if re.search(r"\bFFMA\b", l):
            assert re.search(r", 128, |12582912|1\.0863247", l), l


def test_small_block_estimator_and_latency_shaped_clock_recovery(sass):
    """k_nest2 (three 200-thread blocks per SM): TMA-fed (UBLKCP + mbarrier phase checks), packed FFMA2 tap loop -- 2 chunks
    x 8 steps x 16 accumulators, twice over (register window offset 0 and 8) -- and within the register budget of three
    resident blocks.  Its resume role and the stand-alone clock-recovery kernel read the interpolator taps with two
    16-byte shared loads per step and tie the upper taps to the lower ones with a byte permute; the resume copies
    16-byte groups from the channel-major demod copy (LDGSTS.128)."""
    nest2 = body(sass, "k_nest2ILi4ELi100")
    assert any("UBLKCP" in l for l in nest2) and any("SYNCS.PHASECHK" in l for l in nest2)
    assert sum(" FFMA2 " in l for l in nest2) >= 2 * 8 * 16
    assert sum("LDGSTS.E.128" in l for l in nest2) >= 2
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    m = re.search(r"k_nest2ILi4ELi100[^\n]*\n\s*REG:(\d+)", res)
    assert m and int(m.group(1)) * 200 * 3 <= 65536, m and m.group(1)
    mm = body(sass, "k_mm_stateless_v2ILi64")
    assert sum("LDS.128" in l for l in mm) >= 16 and sum(" PRMT " in l for l in mm) >= 32
    # one rounding per operation survives the load scheduling: 8 FMUL + 8 FADD per step, no FFMA in the interpolation
    assert sum(" FMUL " in l for l in mm) >= 64 and sum(" FADD " in l for l in mm) >= 64

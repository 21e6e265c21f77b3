/* oracle/qts_oracle.c -- TEST INFRASTRUCTURE (see oracle.h).  CPU restatement of the reference's lossy signal degradation,
 * the per-sample step of `slow5tools degrade`:
 *   round_to_power_of_2   slow5lib/src/slow5_press.c:1965-1985   zero the b low bits of an int by rounding to the nearest
 *                                                                 multiple of 2^b (the dropped bits >= 2^(b-1) round up)
 *   slow5_arr_qts_round   slow5lib/src/slow5_press.c:1991-2005   that, for every sample, result stored back as int16; b = 0 is a no-op
 * Written as the reference computes it (int arithmetic on the sign-extended sample, two's complement masks) so the wrap at the
 * int16 store is the reference's.
 * Parity pinning (tests/test_degrade.py): the reference's golden pair test/data/raw/degrade/example2.slow5 ->
 * test/data/exp/degrade/example2_b1.slow5 (test/test_degrade.sh testcase 1), the compiled reference's slow5_arr_qts_round on
 * random arrays for b = 1..16 when oracle/_ref is present, and the signals decoded from the b3 / b4 BLOW5 goldens. */
#include "oracle.h"

int orc_qts_round_sample(int sample, int bits) {
    const int dropped_mask = (1 << bits) - 1;
    const int dropped = sample & dropped_mask;
    const int kept = sample & ~dropped_mask;
    return dropped >= (1 << (bits - 1)) ? kept + (1 << bits) : kept;
}

void orc_qts_round(int16_t *samples, uint64_t n, int bits) {
    if (bits == 0) return;
    for (uint64_t i = 0; i < n; ++i) samples[i] = (int16_t)orc_qts_round_sample(samples[i], bits);
}

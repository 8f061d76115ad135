/* ORACLE (test infrastructure): BabyBear root-of-unity tables. Canonical values of ROU_FWD from
 * risc0-core 3.0.0 field/baby_bear.rs as reproduced and verified in SURVEY.md Appendix B.1
 * (ROU_FWD[k] = 137^(2^(27-k)) mod P). */
#include "bb.h"

const uint32_t BB_ROU_FWD_CANON[BB_MAX_ROU_PO2 + 1] = {
    1,          2013265920, 284861408,  1801542727, 567209306,  740045640,  918899846,
    1881002012, 1453957774, 65325759,   1538055801, 515192888,  483885487,  157393079,
    1695124103, 2005211659, 1540072241, 88064245,   1542985445, 1269900459, 1461624142,
    825701067,  682402162,  1311873874, 1164520853, 352275361,  18769,      137};

fp bb_rou_fwd(unsigned po2) { return fp_from_u32(BB_ROU_FWD_CANON[po2]); }
fp bb_rou_rev(unsigned po2) { return fp_inv(bb_rou_fwd(po2)); }

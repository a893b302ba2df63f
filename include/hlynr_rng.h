/*
 * hlynr_rng.h -- the site-keyed random-draw contract (constants only, no code).
 *
 * The reference draws from four independent NumPy generators in data-dependent order
 * (SURVEY 8a "RNG site table": rl_system/environment.py:409-467,1105,1128; core.py:417-426,
 * 470,563; physics_models.py:372-384; physics_randomizer.py:166-214).  A sequential stream
 * cannot be reproduced by a batched kernel, so every draw site gets a fixed address in a
 * counter-based generator instead; a site's draw is always defined and merely unused when the
 * reference would not have drawn.
 *
 *   generator  Philox4x32-10 (Salmon et al., SC'11; Random123 v1.09 constants)
 *   key        (seed & 0xffffffff, seed >> 32)
 *   counter    (global_env_id & 0xffffffff, episode, step, block | (global_env_id >> 32) << 16)
 *                 episode = number of resets of that env so far minus one (first episode 0)
 *                 step    = 0 for the draws made inside reset(), else env.steps after increment
 *   output     4 x uint32 per block: x0..x3
 *
 *   uniform  [0,1):  u(x)  = (x >> 8) * 2^-24                      (exact in float32)
 *   uniform  (0,1]:  uo(x) = ((x >> 8) + 1) * 2^-24
 *   normals  (float32 Box-Muller, 4 per block):
 *        r = sqrtf(-2 logf(uo(x0))), t = 2 * u(x1) (exact):  z0 = r cospif(t), z1 = r sinpif(t)
 *        r = sqrtf(-2 logf(uo(x2))), t = 2 * u(x3) (exact):  z2 = r cospif(t), z3 = r sinpif(t)
 *        (cospif(t) = cos(pi t) evaluated to float accuracy; no range reduction is needed for t in [0,2))
 *   exponential(1):  e = -logf(uo(x0))
 *
 * The same draws are injected into the unmodified reference by oracle/ref_harness.py (tape
 * objects replacing the four NumPy generators), which is how "identical injected noise draws"
 * (north_star) is realised.
 */
#ifndef HLYNR_RNG_H
#define HLYNR_RNG_H

#define HLYNR_PHILOX_M0 0xD2511F53u
#define HLYNR_PHILOX_M1 0xCD9E8D57u
#define HLYNR_PHILOX_W0 0x9E3779B9u
#define HLYNR_PHILOX_W1 0xBB67AE85u

/* per-tick blocks */
#define HLYNR_BLK_EVADE 0u    /* normals z0..z2: missile evasion, environment.py:1105 */
#define HLYNR_BLK_WIND 1u     /* normals z0..z2: turbulence (physics_models.py:372) or AR(1) wind (environment.py:1128) */
#define HLYNR_BLK_UNI 2u      /* uniforms: x0 gust (physics_models.py:381), x1 onboard dropout (core.py:563),
                                 x2 ground dropout (core.py:417), x3 datalink packet loss (core.py:470) */
#define HLYNR_BLK_GPOS 3u     /* normals z0..z2: ground position noise, core.py:425 */
#define HLYNR_BLK_GVEL 4u     /* normals z0..z2: ground velocity noise, core.py:426 */
#define HLYNR_BLK_GUST_DIR 5u /* normals z0..z2: gust direction, physics_models.py:382 */
#define HLYNR_BLK_GUST_MAG 6u /* exponential from x0: gust magnitude, physics_models.py:384 */
/* reset blocks (step = 0) */
#define HLYNR_BLK_SPAWN0 8u   /* uniforms: missile position x,y,z (or radius, azimuth, elevation), missile speed */
#define HLYNR_BLK_SPAWN1 9u   /* uniforms: interceptor position x,y,z, interceptor velocity x (or speed) */
#define HLYNR_BLK_SPAWN2 10u  /* uniforms: interceptor velocity y,z */
#define HLYNR_BLK_DR0 12u     /* normals: DR draws 0..3  (physics_randomizer.py:166-214 order) */
#define HLYNR_BLK_DR1 13u     /* DR draws 4..7 */
#define HLYNR_BLK_DR2 14u     /* DR draws 8..11 */
#define HLYNR_BLK_DR3 15u     /* DR draw 12 */
/* synthetic random policy (hlynr_rollout with actions == NULL): a = 2u-1 */
#define HLYNR_BLK_ACT0 16u    /* a0..a3 */
#define HLYNR_BLK_ACT1 17u    /* a4..a5 */

/* volley mode (environment.py:42-44): missile m >= 1 of an env draws from its own blocks (missile 0 uses
 * HLYNR_BLK_SPAWN0 / HLYNR_BLK_EVADE, so a volley of one missile draws exactly like the single-missile mode) */
#define HLYNR_BLK_VSPAWN(m) (32u + (unsigned)(m)) /* uniforms: position x,y,z (or radius, azimuth, elevation), speed of missile m */
#define HLYNR_BLK_VEVADE(m) (48u + (unsigned)(m)) /* normals z0..z2: evasion of missile m */

#endif

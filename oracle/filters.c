/*
 * filters.c -- oracle restatement of the servo loop's signal filters.
 *
 * TEST INFRASTRUCTURE (see go1_oracle.h).  Restates (GO1 = unitree_ros/go1_rt_control):
 *   butterworthLPF::init / filter      GO1/src/Filter/butterworthLPF.cpp:82-121   second-order Butterworth low-pass with a
 *                                      first-order start-up for its first three samples; go1_servo runs 28 of them on the
 *                                      slots of the /MPC/Gait message (servo.cpp:579-610, 898-931)
 *   ButterworthFilter::ForceFilter     GO1/src/Filter/butterworth_filter.cpp:37-69 fixed-coefficient force filter
 * Pinned bit for bit against the unmodified classes (oracle/_ref/libref_filter.so): tests/golden/filter_ref.npz + live.
 * Frozen quirks: pi is the literal 3.14159265359; the start-up branch runs while the call counter is <= 2 and the counter
 * stops at 3; the force filter's b[0] is 0 (the current raw sample never enters).
 */
#include <math.h>
#include "go1_oracle.h"

void orc_lpf_init(double fsampling, double fcutoff, orc_lpf_coef *c)
{
    const double ff = fcutoff / fsampling;
    const double ita = 1.0 / tan(3.14159265359 * ff);
    const double q = sqrt(2.0);
    c->b0 = 1.0 / (1.0 + q * ita + ita * ita);
    c->b1 = 2 * c->b0;
    c->b2 = c->b0;
    c->a1 = 2.0 * (ita * ita - 1.0) * c->b0;
    c->a2 = -(1.0 - q * ita + ita * ita) * c->b0;
    c->a = (2.0 * 3.14159265359 * ff) / (2.0 * 3.14159265359 * ff + 1.0);
}

/* state: [0] call counter i, [1] y_p, [2] y_pp, [3] x_p, [4] x_pp */
double orc_lpf_filter(const orc_lpf_coef *c, double s[5], double y)
{
    double out;
    if (s[0] > 2)
        out = c->b0 * y + c->b1 * s[1] + c->b2 * s[2] + c->a1 * s[3] + c->a2 * s[4];
    else {
        out = s[3] + c->a * (y - s[3]);
        s[0] += 1;
    }
    s[2] = s[1]; s[1] = y; s[4] = s[3]; s[3] = out;
    return out;
}

/* state: [0] count, [1..2] raw data, [3..5] filtered data */
double orc_force_filter(double s[6], double input)
{
    const double a[3] = { 1.0, -1.6498, 0.7022 };
    const double b[2] = { 0.0, 0.0521 };
    if (s[0] == 0) {
        s[5] = input; s[2] = input; s[0] += 1;
    } else if (s[0] == 1) {
        s[4] = s[5]; s[5] = input;
        s[1] = s[2]; s[2] = input;
        s[0] += 1;
    } else {
        s[1] = s[2]; s[2] = input;
        s[3] = s[4]; s[4] = s[5];
        s[5] = b[0] * s[2] + b[1] * s[1] - a[1] * s[4] - a[2] * s[3];
    }
    return s[5];
}

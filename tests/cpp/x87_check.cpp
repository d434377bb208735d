// Host check of crypto_recommendation_b200/csrc/x87.cuh (the product's emulation of the reference's x87 arithmetic)
// against real `long double` code written the way cust_vector.hpp:107-174 is: products rounded to double, accumulated
// in long double, quotient formed in long double, result converted to double.  Prints the number of mismatches.
#include <cstdio>
#include <random>
#include <vector>

#include "x87.cuh"

int main(int argc, char** argv) {
    long trials = argc > 1 ? atol(argv[1]) : 200000;
    std::mt19937_64 g(7);
    std::normal_distribution<double> nd(0, 1);
    std::uniform_real_distribution<double> ud(0, 1);
    long bad_dot = 0, bad_cos = 0, plain_differs = 0;
    for (long trial = 0; trial < trials; trial++) {
        int D = 1 + (int)(ud(g) * 203);
        std::vector<double> a(D), b(D);
        int mode = (int)(trial % 5);
        for (int i = 0; i < D; i++) {
            if (mode == 0) { a[i] = nd(g); b[i] = nd(g); }
            else if (mode == 1) { a[i] = (float)nd(g); b[i] = (float)nd(g); }
            else if (mode == 2) { a[i] = 0.37; b[i] = 0.52; if (ud(g) < 0.07) a[i] += ud(g); if (ud(g) < 0.07) b[i] += ud(g); }
            else if (mode == 3) { a[i] = nd(g) * std::ldexp(1.0, (int)(ud(g) * 60) - 30); b[i] = nd(g) * std::ldexp(1.0, (int)(ud(g) * 60) - 30); }
            else { a[i] = (i % 3 == 0) ? nd(g) : 0.0; b[i] = (i % 2 == 0) ? -a[i] : nd(g); }   // cancellation, zeros
        }
        long double acc = 0.0L;
        double na = 0, nb = 0;
        X87 e = {0.0, 0.0};
        for (int i = 0; i < D; i++) {
            double p = a[i] * b[i];
            acc = acc + p;
            x87_add(e, p);
            na = na + a[i] * a[i];
            nb = nb + b[i] * b[i];
        }
        if ((long double)e.h + (long double)e.l != acc) bad_dot++;
        if (x87_to_double(e) != (double)acc) bad_dot++;
        double denom = std::sqrt(na) * std::sqrt(nb);
        double ref = (double)(acc / denom);
        double got = cos_sim_x87(e, na, nb);
        if (!(ref == got) && !(ref != ref && got != got)) bad_cos++;
        if (ref != (double)acc / denom) plain_differs++;
    }
    std::printf("trials %ld dot_mismatches %ld cos_mismatches %ld plain_double_differs %ld\n", trials, bad_dot, bad_cos, plain_differs);
    return (bad_dot || bad_cos) ? 1 : 0;
}

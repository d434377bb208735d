/*
 * main_ref_wrap.cpp -- builds the reference's OWN main.cpp into oracle/_ref/recommendation_ref.
 * TEST INFRASTRUCTURE ONLY.  No algorithm here: the translation unit is the unmodified main.cpp (resolved through
 * -I$(CRX_REF_DIR)) plus the two things needed to run it at all (see ref_harness.cpp): the explicit specialisation
 * of CustHashtable<double>::insertVector that returns a value, and a settable clock behind the token
 * `system_clock` so the RNG seeds are reproducible (seed = environment variable CRX_FAKE_SEED, default 1); the token
 * `time` is re-pointed the same way for the srand((int)time(0)) calls of the validation helpers.  Both builds (reference
 * headers / drop-in headers) get exactly these two substitutions and nothing else.
 */
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <ctime>
#include <fstream>
#include <functional>
#include <iostream>
#include <random>
#include <set>
#include <sstream>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

namespace std { namespace chrono {
struct crx_fake_clock {
    struct dur { unsigned long count() const { const char* s = getenv("CRX_FAKE_SEED"); return s ? strtoul(s, nullptr, 10) : 1ul; } };
    struct tp { dur time_since_epoch() const { return dur(); } };
    static tp now() { return tp(); }
};
} }
#define system_clock crx_fake_clock

/* the 10-fold validation helpers reseed rand() with srand((int)time(0)) (crypto_rec.hpp:350,410): pin that clock too */
static time_t crx_fake_time(time_t*) { const char* s = getenv("CRX_FAKE_SEED"); return (time_t)(s ? strtoul(s, nullptr, 10) : 1ul); }
#define time crx_fake_time

#ifndef CRX_DROPIN_BUILD
/* reference build: its headers, in main.cpp's include order, then the insertVector fix */
#include "lib/in_out/arg_parser.h"
#include "lib/in_out/vector_reader.hpp"
#include "lib/data_structures/cust_vector.hpp"
#include "lib/data_structures/cust_hashtable.hpp"
template <>
int CustHashtable<double>::insertVector(CustVector<double>* inVector) {
    unsigned int index = mod(hashGenerator->generate(inVector), buckets.size());
    buckets[index]->insertVector(inVector);
    return 0;
}
#endif

#include "main.cpp"

// Drop-in seat of the reference's engine/ivf_flat_index.h: a reference call site that says
// #include "../engine/ivf_flat_index.h" (test/simple_test.cpp:7, bench/benchmark.cpp:14) gets the B200 mirror.
#pragma once
#include "../../ivf_flat_index.h"

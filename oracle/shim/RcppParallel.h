// TEST INFRASTRUCTURE ONLY (oracle/): empty stand-in for "RcppParallel.h"
// (included by the reference's src/chain.cpp:11 only to pull in TBB for <execution>).
#pragma once

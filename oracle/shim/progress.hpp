// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for RcppProgress <progress.hpp>.
#pragma once
#include "progress_bar.hpp"
class Progress
{
   public:
    Progress(unsigned long, bool) {}
    Progress(unsigned long, bool, ProgressBar &) {}
    bool increment(unsigned long = 1) { return true; }
    static bool check_abort() { return false; }
};

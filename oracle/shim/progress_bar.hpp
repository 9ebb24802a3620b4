// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for RcppProgress <progress_bar.hpp>.
#pragma once
class ProgressBar
{
   public:
    virtual ~ProgressBar() {}
    virtual void display() {}
    virtual void update(float) {}
    virtual void end_display() {}
};

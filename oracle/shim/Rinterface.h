// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for <Rinterface.h>.
#pragma once
inline void R_FlushConsole() {}

// TEST INFRASTRUCTURE ONLY (oracle/): empty stand-in for <R.h>.
#pragma once

// TEST INFRASTRUCTURE ONLY — see core.hpp in this directory.
#include "core.hpp"

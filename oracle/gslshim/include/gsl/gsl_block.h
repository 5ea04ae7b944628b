/* stub: forwards to the from-scratch GSL-subset shim (test infrastructure, not GSL) */
#include "gslshim.h"

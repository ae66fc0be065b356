// included by the reference's src/core/lib/lattice/lattice.cpp:53 (through DCRTPOLY_IMPLEMENTATION) right before the
// explicit instantiations: member definitions of the host mirror class, then of the B200 class
#include "lattice/hal/default/dcrtpoly-impl.h"
#include "lattice/hal/b200/dcrtpoly-b200-impl.h"

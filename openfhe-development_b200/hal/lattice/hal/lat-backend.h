// lat-backend.h -- B200 variant of the reference's lattice backend selection header
// (src/core/include/lattice/hal/lat-backend.h:39-100).  It is found FIRST on the include path when the reference is
// built with -I openfhe-development_b200/hal (oracle/Makefile.hal) and differs from the reference's in one respect:
// the DCRTPoly aliases name lbcrypto::b200::DCRTPolyB200 (lattice/hal/b200/dcrtpoly-b200.h), whose hot methods run on
// the GPU through libofhe_b200.so.  Parameter classes, Poly / NativePoly and the default DCRTPolyImpl (kept as the
// backend's host mirror type) are the reference's own, included from where they lie.
#ifndef LBCRYPTO_INC_LATTICE_HAL_LAT_BACKEND_H
#define LBCRYPTO_INC_LATTICE_HAL_LAT_BACKEND_H

#define ILPARAMS_IMPLEMENTATION     "lattice/hal/default/ilparams.h"
#define ILDCRTPARAMS_IMPLEMENTATION "lattice/hal/default/ildcrtparams.h"
#define POLY_IMPLEMENTATION         "lattice/hal/default/poly-impl.h"
// lattice.cpp includes this for the explicit instantiations: the mirror's member definitions, then ours
#define DCRTPOLY_IMPLEMENTATION     "lattice/hal/b200/dcrtpoly-b200-instantiate.h"

#define MAKE_ILPARAMS_TYPE(T)     template class ILParamsImpl<T>;
#define MAKE_ILDCRTPARAMS_TYPE(T) template class ILDCRTParams<T>;
#define MAKE_POLY_TYPE(T)         template class PolyImpl<T>;
#define MAKE_DCRTPOLY_TYPE(T)     \
    template class DCRTPolyImpl<T>; \
    template class b200::DCRTPolyB200<T>;

#include "lattice/hal/default/ilparams.h"
#include "lattice/hal/default/ildcrtparams.h"
#include "lattice/hal/default/poly.h"
#include "lattice/hal/default/dcrtpoly.h"
#include "lattice/hal/b200/dcrtpoly-b200.h"

namespace lbcrypto {

using ILNativeParams = ILParamsImpl<NativeInteger>;
using ILParams       = ILParamsImpl<BigInteger>;
using Poly           = PolyImpl<BigVector>;
using NativePoly     = PolyImpl<NativeVector>;
using DCRTPoly       = b200::DCRTPolyB200<BigVector>;

#ifdef WITH_BE2
using M2Params     = ILParamsImpl<M2Integer>;
using M2DCRTParams = ILDCRTParams<M2Integer>;
using M2Poly       = PolyImpl<M2Vector>;
using M2DCRTPoly   = b200::DCRTPolyB200<M2Vector>;
#else
using M2Params     = void;
using M2DCRTParams = void;
using M2Poly       = void;
using M2DCRTPoly   = void;
#endif

#ifdef WITH_BE4
using M4Params     = ILParamsImpl<M4Integer>;
using M4DCRTParams = ILDCRTParams<M4Integer>;
using M4Poly       = PolyImpl<M4Vector>;
using M4DCRTPoly   = b200::DCRTPolyB200<M4Vector>;
#else
using M4Params     = void;
using M4DCRTParams = void;
using M4Poly       = void;
using M4DCRTPoly   = void;
#endif

#ifdef WITH_NTL
using M6Params     = ILParamsImpl<M6Integer>;
using M6DCRTParams = ILDCRTParams<M6Integer>;
using M6Poly       = PolyImpl<M6Vector>;
using M6DCRTPoly   = b200::DCRTPolyB200<M6Vector>;
#else
using M6Params     = void;
using M6DCRTParams = void;
using M6Poly       = void;
using M6DCRTPoly   = void;
#endif

}  // namespace lbcrypto

#endif

//========================================================================================
// TEST INFRASTRUCTURE ONLY (oracle/): a problem generator written for this repository and
// compiled against the UNMODIFIED reference by oracle/build_ref.py.  It exists to pin the
// reference's user-hook paths that none of its own simple pgens exercise:
//   * Mesh::EnrollUserExplicitSourceFunction (SrcTermFunc, src/athena.hpp:185-189), called
//     last in HydroSourceTerms::AddSourceTerms (hydro/srcterms/hydro_srcterms.cpp:150-153);
// with arithmetic restricted to + - * / sqrt so that tests/util.py can restate the very
// same function in numpy, bit for bit.
//========================================================================================
#include <cmath>

#include "../athena.hpp"
#include "../athena_arrays.hpp"
#include "../coordinates/coordinates.hpp"
#include "../eos/eos.hpp"
#include "../field/field.hpp"
#include "../hydro/hydro.hpp"
#include "../mesh/mesh.hpp"
#include "../parameter_input.hpp"
#include "../scalars/scalars.hpp"

namespace {
Real gm, soft2, sdecay;
// softened point-mass gravity whose strength grows linearly in time, plus scalar decay
void CentralGravity(MeshBlock *pmb, const Real time, const Real dt,
                    const AthenaArray<Real> &prim, const AthenaArray<Real> &prim_scalar,
                    const AthenaArray<Real> &bcc, AthenaArray<Real> &cons,
                    AthenaArray<Real> &cons_scalar) {
  Real amp = gm*(1.0 + 0.5*time);
  for (int k=pmb->ks; k<=pmb->ke; ++k) {
    Real z = pmb->pcoord->x3v(k);
    for (int j=pmb->js; j<=pmb->je; ++j) {
      Real y = pmb->pcoord->x2v(j);
      for (int i=pmb->is; i<=pmb->ie; ++i) {
        Real x = pmb->pcoord->x1v(i);
        Real rsq = (x*x + y*y) + (z*z + soft2);
        Real r = std::sqrt(rsq);
        Real fac = amp/(rsq*r);
        Real den = prim(IDN,k,j,i);
        Real s1 = (dt*den)*(fac*x);
        Real s2 = (dt*den)*(fac*y);
        Real s3 = (dt*den)*(fac*z);
        cons(IM1,k,j,i) -= s1;
        cons(IM2,k,j,i) -= s2;
        cons(IM3,k,j,i) -= s3;
        if (NON_BAROTROPIC_EOS) {
          cons(IEN,k,j,i) -= (s1*prim(IVX,k,j,i) + s2*prim(IVY,k,j,i)) + s3*prim(IVZ,k,j,i);
        }
        for (int n=0; n<NSCALARS; ++n) {
          cons_scalar(n,k,j,i) -= (dt*sdecay)*(den*prim_scalar(n,k,j,i));
        }
      }
    }
  }
}
} // namespace

void Mesh::InitUserMeshData(ParameterInput *pin) {
  gm = pin->GetOrAddReal("problem", "gm", 0.5);
  soft2 = pin->GetOrAddReal("problem", "soft2", 0.01);
  sdecay = pin->GetOrAddReal("problem", "sdecay", 0.3);
  EnrollUserExplicitSourceFunction(CentralGravity);
}

// a smooth, non-symmetric state built from the cell-centre coordinates (rational functions)
void MeshBlock::ProblemGenerator(ParameterInput *pin) {
  Real gam = NON_BAROTROPIC_EOS ? pin->GetReal("hydro", "gamma") : 0.0;
  for (int k=ks; k<=ke; ++k) {
    for (int j=js; j<=je; ++j) {
      for (int i=is; i<=ie; ++i) {
        Real x = pcoord->x1v(i), y = pcoord->x2v(j), z = pcoord->x3v(k);
        Real d = 1.0 + 0.5/(1.0 + 8.0*((x-0.1)*(x-0.1) + (y+0.2)*(y+0.2) + z*z));
        phydro->u(IDN,k,j,i) = d;
        phydro->u(IM1,k,j,i) = d*(0.3*y - 0.1*z);
        phydro->u(IM2,k,j,i) = d*(-0.3*x + 0.2*z);
        phydro->u(IM3,k,j,i) = d*(0.1*x*y);
        if (NON_BAROTROPIC_EOS) {
          Real p = 0.6 + 0.2*x*x;
          phydro->u(IEN,k,j,i) = p/(gam - 1.0) + 0.5*(SQR(phydro->u(IM1,k,j,i))
              + SQR(phydro->u(IM2,k,j,i)) + SQR(phydro->u(IM3,k,j,i)))/d;
        }
        for (int n=0; n<NSCALARS; ++n)
          pscalars->s(n,k,j,i) = d*(0.5 + 0.4*x/(1.0 + n))/(1.0 + y*y);
      }
    }
  }
  if (MAGNETIC_FIELDS_ENABLED) {
    Real b1 = 0.4, b2 = -0.3, b3 = 0.2;     // uniform field: divergence-free on the faces
    for (int k=ks; k<=ke; ++k) for (int j=js; j<=je; ++j) for (int i=is; i<=ie+1; ++i)
      pfield->b.x1f(k,j,i) = b1;
    for (int k=ks; k<=ke; ++k) for (int j=js; j<=je+1; ++j) for (int i=is; i<=ie; ++i)
      pfield->b.x2f(k,j,i) = b2;
    for (int k=ks; k<=ke+1; ++k) for (int j=js; j<=je; ++j) for (int i=is; i<=ie; ++i)
      pfield->b.x3f(k,j,i) = b3;
    if (NON_BAROTROPIC_EOS) {
      for (int k=ks; k<=ke; ++k) for (int j=js; j<=je; ++j) for (int i=is; i<=ie; ++i)
        phydro->u(IEN,k,j,i) += 0.5*(b1*b1 + b2*b2 + b3*b3);
    }
  }
}

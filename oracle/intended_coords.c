/* intended_coords.c -- OUR definition of GetElementCoords with the semantics the reference's author commented
 * out (src/Discretization.c:40-43): node0=(ei,ej), node1=(ei,ej+1), node2=(ei+1,ej+1), node3=(ei+1,ej).
 * The live lines :34-38 give all four nodes the coordinate of node (ei,ej) (singular Jacobian, NaN matrix;
 * SURVEY.md Appendix B item 1).  This file is NOT copied from the reference: it is loaded BEFORE the
 * unmodified reference object (symbol interposition: the reference is compiled -fPIC, so its own call to
 * GetElementCoords goes through the PLT) to run the reference's assembly loops in "intended" mode without
 * touching its sources. */
#include <petsc.h>

PetscErrorCode GetElementCoords(DMDACoor2d **_coords, PetscInt ei, PetscInt ej, PetscScalar *el_coords) {
  el_coords[0] = _coords[ej][ei].x;         el_coords[1] = _coords[ej][ei].y;
  el_coords[2] = _coords[ej + 1][ei].x;     el_coords[3] = _coords[ej + 1][ei].y;
  el_coords[4] = _coords[ej + 1][ei + 1].x; el_coords[5] = _coords[ej + 1][ei + 1].y;
  el_coords[6] = _coords[ej][ei + 1].x;     el_coords[7] = _coords[ej][ei + 1].y;
  return 0;
}

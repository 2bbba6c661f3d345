"""Problem generators: the MeshBlock::ProblemGenerator(pin) hook of the reference
(src/pgen/default_pgen.cpp).  Each generator is a function `pgen(pmb, pin) -> dict` returning
the conserved variables `u` (active cells filled, AthenaArray layout incl. ghosts) and, with
MHD, the face fields `b1,b2,b3`, exactly what the reference's pgens write into phydro->u and
pfield->b.  They run on the host (as in the reference); Mesh.problem_generator uploads."""
from .blast import blast  # noqa: F401
from .linear_wave import linear_wave, linear_wave_errors  # noqa: F401
from .orszag_tang import orszag_tang  # noqa: F401
from .kh import kh  # noqa: F401
from .shock_tube import shock_tube  # noqa: F401

BY_NAME = {"blast": blast, "linear_wave": linear_wave, "orszag_tang": orszag_tang, "kh": kh,
           "shock_tube": shock_tube}

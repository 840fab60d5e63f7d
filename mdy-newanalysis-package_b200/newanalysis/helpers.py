"""newanalysis.helpers (B200 backend): only the producers of the g-function path's inputs live here.

comByResidue, dipByResidue and velcomByResidue keep the reference's signatures
(/root/reference/newanalysis/helpers/helpers.pyx:41-123) and return bit-identical arrays; ResidueMap keeps
the topology on the device and can hand device-resident centres of mass / dipoles straight to
newanalysis.gfunction.RDF (no host round trip).  The other ~50 helpers of the reference module are out of
scope.
"""
from .gfunction import ResidueMap, comByResidue, dipByResidue, velcomByResidue  # noqa: F401

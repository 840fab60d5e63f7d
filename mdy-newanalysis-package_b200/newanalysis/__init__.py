"""newanalysis (B200 backend): only the g-function pair path lives here.

`from newanalysis.gfunction import RDF` is the reference's import (docs/notebooks/rdf.ipynb); the
extension module links libgfn.so that sits next to it.  Importing this package never imports torch.
"""

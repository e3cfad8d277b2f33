"""Stand-in for the absent, unpinned `geomloss` package (test infrastructure
only; see oracle/sinkhorn.py for the restated algorithm and the
"parity unpinned" note)."""
from oracle.sinkhorn import SamplesLoss  # noqa: F401

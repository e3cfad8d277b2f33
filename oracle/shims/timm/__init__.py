"""Stub of the absent `timm` package (oracle/test infrastructure only).

`/root/reference/model/models.py:6` does `import timm` and calls
`timm.create_model` (models.py:60-68).  timm==0.9.12 is pinned in the
reference's requirements.txt:28 but is not installed in this image, so the
golden-vector generator puts this directory ahead of `/root/reference` on
sys.path.  Nothing under deltakd_b200/ imports it.
"""


def create_model(*args, **kwargs):  # replaced by the generator with stand-ins
    raise RuntimeError("timm is not installed; oracle/make_golden.py patches this")

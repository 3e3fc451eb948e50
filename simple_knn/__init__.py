"""Import-compatible alias of the third-party ``simple_knn`` package (``from simple_knn._C import distCUDA2``,
/root/reference/core/gaussians/gs.py:6)."""

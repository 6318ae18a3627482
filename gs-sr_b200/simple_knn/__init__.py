"""Drop-in replacement for GS-SR's ``simple_knn`` extension; see ``simple_knn._C.distCUDA2``."""

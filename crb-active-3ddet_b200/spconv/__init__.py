"""Drop-in `spconv` package backed by libcrb3d_sm100 (see spconv/pytorch and spconv/utils)."""
__version__ = "2.1.21+crb3d.b200"

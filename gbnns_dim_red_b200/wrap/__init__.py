"""Drop-in for the reference's `wrap` namespace package (SWIG module `wrap.c_support`)."""

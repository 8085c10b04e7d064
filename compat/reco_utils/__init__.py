"""Import-path shim: see compat/README.md."""
__version__ = "pamrec_b200-shim"

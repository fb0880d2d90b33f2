"""`models` package name of the reference checkout, resolved to the B200 engine (see compat/README.md)."""

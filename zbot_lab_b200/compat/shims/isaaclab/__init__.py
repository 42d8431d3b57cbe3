"""shim: see ../README.md"""

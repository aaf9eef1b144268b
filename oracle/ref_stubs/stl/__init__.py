class mesh:  # numpy-stl stand-in; only used by export_stl (out of scope)
    pass

def typechecked(f=None, **kw):
    if f is None:
        return lambda g: g
    return f

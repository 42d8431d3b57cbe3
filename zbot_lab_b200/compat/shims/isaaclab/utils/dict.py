def print_dict(val, nesting: int = -4, start: bool = True):
    """``isaaclab.utils.dict.print_dict`` (train.py:90)."""
    if isinstance(val, dict):
        if not start:
            print("")
        nesting += 4
        for k in val:
            print(nesting * " ", end="")
            print(k, end=": ")
            print_dict(val[k], nesting, start=False)
    else:
        print(val)

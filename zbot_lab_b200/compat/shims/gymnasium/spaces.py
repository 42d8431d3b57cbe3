def flatdim(space) -> int:
    if isinstance(space, int):
        return space
    n = 1
    for s in space.shape:
        n *= int(s)
    return n

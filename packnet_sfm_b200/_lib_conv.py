"""ctypes declarations for the convolution / layer entry points (filled in as they land)."""


def declare(lib):
    return lib

from _dummy import Dummy


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    return Dummy()

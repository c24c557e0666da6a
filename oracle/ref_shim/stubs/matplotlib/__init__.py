"""Empty stand-in so `import kbbq` (reference) succeeds without matplotlib. Test infrastructure only."""


def use(*args, **kwargs):
    pass

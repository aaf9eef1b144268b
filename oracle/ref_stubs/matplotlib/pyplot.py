"""Stand-in for matplotlib.pyplot: the reference imports it at module scope but the solve path never calls it."""


def __getattr__(name):
    raise RuntimeError("matplotlib stub: pyplot.%s is not available in the oracle harness" % name)

"""Import stub so that the reference's qpth/env_dx/*.py (which plot in methods the goldens never call) can be
imported in the build container, where matplotlib is not installed.  TEST INFRASTRUCTURE, NOT PRODUCT."""


def use(*a, **k):
    return None

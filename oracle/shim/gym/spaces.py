class _Space(object):
    def __init__(self, *a, **k):
        pass


class Discrete(_Space):
    pass


class Box(_Space):
    pass


class Graph(_Space):
    pass

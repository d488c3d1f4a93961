class _Space:
    def __init__(self, *args, **kwargs):
        self.args = args
        self.kwargs = kwargs
        self.shape = kwargs.get("shape")
        self.n = args[0] if args else kwargs.get("n")


class Box(_Space):
    pass


class Discrete(_Space):
    pass


class MultiBinary(_Space):
    pass

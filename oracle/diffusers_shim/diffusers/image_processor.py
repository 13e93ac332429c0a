class VaeImageProcessor:
    def __init__(self, **kw):
        self.kw = kw

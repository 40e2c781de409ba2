def make_dot(*a, **k):
    raise NotImplementedError

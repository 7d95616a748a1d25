"""TEST INFRASTRUCTURE ONLY -- builtin message/reduce descriptors of the dgl shim (see dgl/__init__.py)."""


class _Msg:
    def __init__(self, kind, lhs, rhs, out):
        self.kind, self.lhs, self.rhs, self.out = kind, lhs, rhs, out


class _Red:
    def __init__(self, kind, msg, out):
        self.kind, self.msg, self.out = kind, msg, out


def copy_src(src, out):
    return _Msg('copy_u', src, None, out)


copy_u = copy_src


def u_mul_e(lhs, rhs, out):
    return _Msg('u_mul_e', lhs, rhs, out)


def u_dot_v(lhs, rhs, out):
    return _Msg('u_dot_v', lhs, rhs, out)


def mean(msg, out):
    return _Red('mean', msg, out)


def sum(msg, out):  # noqa: A001 - mirrors dgl.function.sum
    return _Red('sum', msg, out)


def max(msg, out):  # noqa: A001 - mirrors dgl.function.max
    return _Red('max', msg, out)

"""``HParams(**kw)`` with attribute access, ``.values()`` and ``.parse("a=1,b=[..],c={..}")``."""
import ast


def _split_top_level(s):
    parts, depth, cur = [], 0, ""
    for ch in s:
        if ch in "[{(":
            depth += 1
        elif ch in "]})":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        parts.append(cur)
    return parts


class HParams(object):
    def __init__(self, **kwargs):
        self.__dict__.update(kwargs)

    def values(self):
        return dict(self.__dict__)

    def parse(self, s):
        for item in _split_top_level(s or ""):
            if "=" not in item:
                continue
            k, v = item.split("=", 1)
            k, v = k.strip(), v.strip()
            try:
                v = ast.literal_eval(v)
            except Exception:
                pass
            if k not in self.__dict__:
                raise ValueError("Unknown hyperparameter: %s" % k)
            setattr(self, k, v)
        return self

    def __contains__(self, k):
        return k in self.__dict__

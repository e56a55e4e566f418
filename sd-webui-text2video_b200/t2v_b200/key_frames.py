"""Key-frame schedules of the img2vid inpainting weights -- `T2VAnimKeys` / `FrameInterpolater` of the reference
(scripts/t2v_helpers/key_frames.py:9-95), restated without numexpr / pandas (neither is a dependency of this package):
expressions are evaluated by a small whitelisting AST walker over the same variables (`t`, `max_f`, `max_i_f`, `s`).

Behaviour kept, including the reference's quirk: inside `get_inbetweens` the "is this key a plain number" flag is only
refreshed on key frames, so after a key whose value is an EXPRESSION every following un-keyed frame evaluates that same
expression at its own `t` (key_frames.py:33-43).  With the UI default '0:(t/max_i_f), "max_i_f":(1)' this yields the linear
ramp 0 -> 1 over the first `inpainting_frames` frames and 1 afterwards (args.py:232).
"""
import ast
import math
import operator
import re

import numpy as np

_FUNCS = {k: getattr(math, k) for k in ('sin', 'cos', 'tan', 'exp', 'log', 'sqrt', 'floor', 'ceil', 'fabs')}
_FUNCS.update({'abs': abs, 'min': min, 'max': max, 'where': lambda c, a, b: a if c else b})
_BIN = {ast.Add: operator.add, ast.Sub: operator.sub, ast.Mult: operator.mul, ast.Div: operator.truediv, ast.Pow: operator.pow,
        ast.Mod: operator.mod, ast.FloorDiv: operator.floordiv}
_CMP = {ast.Lt: operator.lt, ast.LtE: operator.le, ast.Gt: operator.gt, ast.GtE: operator.ge, ast.Eq: operator.eq,
        ast.NotEq: operator.ne}


def evaluate(expr, variables):
    """Arithmetic expression -> float (the subset of numexpr the schedules use); anything else raises ValueError."""
    def ev(n):
        if isinstance(n, ast.Expression):
            return ev(n.body)
        if isinstance(n, ast.Constant) and isinstance(n.value, (int, float)):
            return n.value
        if isinstance(n, ast.Name):
            if n.id in variables:
                return variables[n.id]
            raise ValueError(f'unknown variable {n.id!r} in key-frame expression {expr!r}')
        if isinstance(n, ast.BinOp) and type(n.op) in _BIN:
            return _BIN[type(n.op)](ev(n.left), ev(n.right))
        if isinstance(n, ast.UnaryOp) and isinstance(n.op, (ast.USub, ast.UAdd)):
            v = ev(n.operand)
            return -v if isinstance(n.op, ast.USub) else v
        if isinstance(n, ast.Compare) and len(n.ops) == 1 and type(n.ops[0]) in _CMP:
            return _CMP[type(n.ops[0])](ev(n.left), ev(n.comparators[0]))
        if isinstance(n, ast.Call) and isinstance(n.func, ast.Name) and n.func.id in _FUNCS and not n.keywords:
            return _FUNCS[n.func.id](*[ev(a) for a in n.args])
        raise ValueError(f'unsupported syntax in key-frame expression {expr!r}')
    return float(ev(ast.parse(expr.strip(), mode='eval')))


def check_is_number(value):
    return re.match(r'^(?=.)([+-]?([0-9]*)(\.([0-9]+))?)$', value)


class FrameInterpolater(object):
    def __init__(self, max_frames=0, seed=-1, max_i_frames=1):
        self.max_frames, self.seed, self.max_i_frames = max_frames, seed, max_i_frames

    def _vars(self, t=0):
        return {'t': t, 'max_f': self.max_frames - 1, 'max_i_f': self.max_i_frames - 1, 's': self.seed}

    @staticmethod
    def sanitize_value(value):
        return value.replace("'", '').replace('"', '').replace('(', '').replace(')', '')

    def parse_key_frames(self, string):
        frames = {}
        for part in string.split(','):
            fp = part.split(':')
            key = fp[0].strip()
            if check_is_number(self.sanitize_value(key)):
                frame = int(self.sanitize_value(key))
            else:
                frame = int(evaluate(key.replace("'", '').replace('"', ''), self._vars()))
            frames[frame] = fp[1].strip()
        if frames == {} and len(string) != 0:
            raise RuntimeError('Key Frame string not correctly formatted')
        return frames

    def get_inbetweens(self, key_frames, integer=False, interp_method='Linear'):
        if interp_method != 'Linear':
            raise NotImplementedError('only the Linear interpolation the inpainting weights use is restated')
        series = np.full(self.max_frames, np.nan, dtype=np.float64)
        value, value_is_number = None, False
        for i in range(self.max_frames):
            if i in key_frames:
                value = key_frames[i]
                value_is_number = bool(check_is_number(self.sanitize_value(value)))
                if value_is_number:
                    series[i] = float(self.sanitize_value(value))
            if value is None:
                raise RuntimeError('the schedule needs a key at frame 0')      # the reference raises NameError here
            if not value_is_number:
                series[i] = evaluate(value, self._vars(t=i))
        valid = np.flatnonzero(~np.isnan(series))
        series[0] = series[valid[0]]
        series[self.max_frames - 1] = series[valid[-1]]
        valid = np.flatnonzero(~np.isnan(series))
        series = np.interp(np.arange(self.max_frames), valid, series[valid])       # pandas interpolate(method='linear', both)
        return series.astype(int) if integer else series


class T2VAnimKeys(object):
    def __init__(self, anim_args, seed=-1, max_i_frames=1):
        self.fi = FrameInterpolater(anim_args.max_frames, seed, max_i_frames)
        self.inpainting_weights_series = self.fi.get_inbetweens(self.fi.parse_key_frames(anim_args.inpainting_weights))

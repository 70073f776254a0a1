"""
``kfunc``: callables with managed keyword parameters -- the interactive front
end the reference puts on its sources and processes (``sdepy.lognorm``,
``sdepy.heston``, ``sdepy.dw`` ...; reference kfun.py:251-356, shortcuts.py:
73-99).  Host-side only: a kfunc'd process re-instantiates the wrapped class
with merged parameters and runs it, so every evaluation goes through the same
CUDA path as the plain classes.

Semantics kept (kfun.py docstring, 262-341):

* parameters are keyword-only and stored in the instance (``.params``);
  variables are the arguments of the wrapped ``__call__`` (positional or by
  name) and are always given at evaluation;
* ``K(**params)`` -> instance; ``K(*vars, **params)`` -> instantiate and
  evaluate at once;
* ``inst(*vars)`` -> evaluate; ``inst(*vars, **params)`` -> evaluate a copy
  with some parameters changed (``inst`` is not affected);
  ``inst(**params)`` -> new instance with merged parameters.
"""
import inspect

__all__ = ['kfunc', 'iskfunc']


def _call_variables(cls):
    """Names of the variables of cls.__call__ (everything but self)."""
    sig = inspect.signature(cls.__call__)
    names = []
    for k, p in list(sig.parameters.items())[1:]:
        if p.kind in (p.VAR_POSITIONAL, p.VAR_KEYWORD):
            continue
        names.append(k)
    return tuple(names)


def _init_defaults(cls):
    out = {}
    for k, p in list(inspect.signature(cls.__init__).parameters.items())[1:]:
        if p.kind == p.KEYWORD_ONLY and p.default is not p.empty:
            out[k] = p.default
    return out


class _kfunc_type(type):
    """Calling the class with variables evaluates the new instance at once."""

    def __call__(cls, *var, **kw):
        variables = {k: kw.pop(k) for k in tuple(kw) if k in cls._kfunc_variables}
        inst = super().__call__(**kw)
        if var or variables:
            return inst._kfunc_evaluate(*var, **variables)
        return inst


def _wrap_class(f):
    if iskfunc(f):
        return f
    if not callable(getattr(f, '__call__', None)) or not inspect.isclass(f):
        raise TypeError('kfunc expects a class with a __call__ method, or '
                        'kfunc(nvar=k) applied to a function')
    base_call = f.__call__

    class wrapper(f, metaclass=_kfunc_type):
        _kfunc_variables = _call_variables(f)
        _kfunc_wrapped = f

        def __init__(self, **params):
            self._kfunc_params = dict(params)
            super().__init__(**params)

        def _kfunc_evaluate(self, *var, **variables):
            return base_call(self, *var, **variables)

        def __call__(self, *var, **kw):
            names = self._kfunc_variables
            variables = {k: kw.pop(k) for k in tuple(kw) if k in names}
            if kw:
                other = type(self)(**{**self._kfunc_params, **kw})
                if not (var or variables):
                    return other
                return other._kfunc_evaluate(*var, **variables)
            return self._kfunc_evaluate(*var, **variables)

        @property
        def params(self):
            """Parameters stored in the instance, constructor defaults and
            (for SDE classes) the SDE-specific ``args`` included."""
            out = _init_defaults(f)
            out.update(getattr(self, 'args', {}) or {})
            out.update(self._kfunc_params)
            return out

    wrapper._kfunc_decorated = True
    for attr in ('__name__', '__qualname__', '__doc__', '__module__'):
        try:
            setattr(wrapper, attr, getattr(f, attr))
        except (AttributeError, TypeError):
            pass
    return wrapper


def _wrap_function(f, nvar):
    """kfunc over a function: the first nvar arguments are variables."""
    names = tuple(inspect.signature(f).parameters)
    variables, accepted = names[:nvar], set(names[nvar:])
    defaults = {k: p.default for k, p in inspect.signature(f).parameters.items()
                if k in accepted and p.default is not p.empty}

    class function_kfunc(metaclass=_kfunc_type):
        _kfunc_variables = variables
        _kfunc_wrapped = f
        _kfunc_decorated = True

        def __init__(self, **params):
            unknown = set(params) - accepted
            if unknown:
                raise TypeError('unexpected keyword(s): {}'.format(unknown))
            self._kfunc_params = dict(params)

        def _kfunc_evaluate(self, *var, **variables):
            return f(*var, **variables, **self._kfunc_params)

        def __call__(self, *var, **kw):
            given = {k: kw.pop(k) for k in tuple(kw) if k in self._kfunc_variables}
            if kw:
                other = type(self)(**{**self._kfunc_params, **kw})
                if not (var or given):
                    return other
                return other._kfunc_evaluate(*var, **given)
            return self._kfunc_evaluate(*var, **given)

        @property
        def params(self):
            return {**defaults, **self._kfunc_params}

    for attr in ('__name__', '__qualname__', '__doc__', '__module__'):
        try:
            setattr(function_kfunc, attr, getattr(f, attr))
        except (AttributeError, TypeError):
            pass
    return function_kfunc


def kfunc(f=None, *, nvar=None):
    """Wrap a class (``kfunc(cls)`` / ``@kfunc``) or a function
    (``@kfunc(nvar=k)``) as a kfunc -- see the module docstring."""
    if f is None:
        if nvar is None:
            raise TypeError('kfunc(nvar=k) expects the number of variables')
        return lambda g: _wrap_function(g, int(nvar))
    if nvar is not None:
        raise TypeError('nvar applies to functions only: use @kfunc(nvar=k)')
    if inspect.isclass(f):
        return _wrap_class(f)
    raise TypeError('to wrap a function use @kfunc(nvar=k)')


def iskfunc(cls_or_object):
    """True for kfunc classes and their instances (reference kfun.py:358)."""
    cls = cls_or_object if inspect.isclass(cls_or_object) else type(cls_or_object)
    return bool(getattr(cls, '_kfunc_decorated', False))

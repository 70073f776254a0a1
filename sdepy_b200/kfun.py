"""
``kfunc``: callables with managed keyword parameters -- the interactive front
end the reference puts on its sources and processes (``sdepy.lognorm``,
``sdepy.heston``, ``sdepy.dw`` ...; reference kfun.py:251-356, shortcuts.py:
73-99).  Host-side only: a kfunc'd process re-instantiates the wrapped class
with merged parameters and runs it, so every evaluation goes through the same
CUDA path as the plain classes.

Semantics kept (reference kfun.py:46-196, 262-356; exercised by the reference's
tests/test_kfunc.py):

* *parameters* are the keyword-only arguments of ``__init__``, stored in the
  instance (``.params``, constructor defaults included; for ``SDE`` classes
  the SDE-specific ``args`` as well); *variables* are the arguments of
  ``__call__`` and are always given at evaluation.  A class whose ``__init__``
  takes anything but keyword-only arguments, whose parameter and variable
  names overlap, that customises ``__new__`` or lacks a user ``__init__`` /
  ``__call__`` is refused with ``TypeError``; an existing ``params`` attribute
  is overwritten with a ``RuntimeWarning``;
* ``K(**params)`` -> instance; ``K(*vars, **params)`` -> instantiate and
  evaluate at once;
* ``inst(*vars)`` -> evaluate; ``inst(*vars, **params)`` -> evaluate a copy
  with some parameters changed (``inst`` is not affected);
  ``inst(**params)`` -> new instance with merged parameters, remembering
  ``inst`` as its ``_kfunc_parent``;
* a subclass of a kfunc class must be decorated again: an undecorated one is
  built and called the plain way, with a ``RuntimeWarning``;
* ``@kfunc(nvar=k)`` wraps a function: its first ``k`` arguments are the
  variables, the rest (keyword-only) the parameters;
* misuse of the decorator (``@kfunc`` on a function, ``@kfunc(nvar=k)`` on a
  class) raises ``SyntaxError``; ``nvar`` out of range ``ValueError``.
"""
import inspect
import warnings

__all__ = ['kfunc', 'iskfunc']

_MARK = '_is_kfunc'
_COPIED = ('__module__', '__name__', '__qualname__', '__doc__')
# keywords this package adds to the reference's constructors: reported by
# ``params`` only when given, so that the parameter sets read like the
# reference's
_EXTENSIONS = ('seed', 'output', 'device', 'path_offset', 'payoff', 'draws')


def _named_like(wrapped, wrapper):
    for attr in _COPIED:
        try:
            setattr(wrapper, attr, getattr(wrapped, attr))
        except (AttributeError, TypeError):
            pass
    return wrapper


def _plain_parameters(func):
    """(name, Parameter) of func's arguments, `self` and **kwargs dropped."""
    items = [(k, p) for k, p in inspect.signature(func).parameters.items()
             if p.kind != p.VAR_KEYWORD]
    return items[1:]


def _defined_by_user(cls, name):
    return any(name in vars(base) for base in cls.__mro__[:-1])


def _sde_base():
    # imported lazily: integration imports the shortcuts that import kfunc
    from .integration import SDE
    return SDE


class _new_only(type):
    """Instantiation goes through __new__ alone, which decides whether the
    result is an instance or the value of an immediate evaluation."""

    def __call__(cls, *var, **kw):
        return cls.__new__(cls, *var, **kw)


def _wrap_class(f):
    if not hasattr(f, _MARK):
        if _defined_by_user(f, '__new__'):
            raise TypeError('class {} customises __new__ and cannot be wrapped '
                            'as a kfunc'.format(f))
        if not (_defined_by_user(f, '__init__') and _defined_by_user(f, '__call__')):
            raise TypeError('cannot wrap {} as a kfunc: user defined __init__ '
                            'and __call__ methods are both needed'.format(f))
        if hasattr(f, 'params'):
            warnings.warn('wrapping {} as a kfunc overwrites its params '
                          'attribute'.format(f), RuntimeWarning)
    init_par = _plain_parameters(f.__init__)
    call_par = _plain_parameters(f.__call__)
    if any(p.kind != p.KEYWORD_ONLY for _, p in init_par):
        raise TypeError('cannot wrap {} as a kfunc: its parameters '
                        '(initialization arguments) should all be '
                        'keyword-only'.format(f))
    defaults = {k: p.default for k, p in init_par}
    variables = dict(call_par)
    if set(defaults) & set(variables):
        raise TypeError('cannot wrap {} as a kfunc: parameters '
                        '(initialization arguments) and variables (calling '
                        'arguments) should not share names'.format(f))
    is_sde = issubclass(f, _sde_base())

    def split(cls, kw):
        # read at call time: _wrap_function re-states the variables afterwards
        names = cls._kfunc_call_args
        var = {k: v for k, v in kw.items() if k in names}
        par = {k: v for k, v in kw.items() if k not in names}
        return var, par

    class wrapper(f, metaclass=_new_only):
        _kfunc_init_args = defaults
        _kfunc_call_args = variables

        def __new__(cls, *var, **kw):
            if _MARK not in vars(cls):
                # undecorated subclass of a kfunc: plain construction
                warnings.warn('a subclass of a kfunc class should be decorated '
                              'with kfunc, but {} was not: it is initialised '
                              'and called as a plain class'.format(cls),
                              RuntimeWarning)
                self = f.__new__(cls)
                self.__init__(*var, **kw)
                return self
            given, params = split(cls, kw)
            self = object.__new__(cls)
            self._kfunc_params = params
            self._kfunc_parent = None
            self.__init__(**params)
            if var or given:
                return super(cls, self).__call__(*var, **given)
            return self

        def __call__(self, *var, **kw):
            given, params = split(type(self), kw)
            if not params:
                return super().__call__(*var, **given)
            # a derived object: stored parameters overridden by the new ones
            merged = {**self._kfunc_params, **params}
            cls = type(self)
            new = object.__new__(cls)
            new.__init__(**merged)
            new._kfunc_params = merged
            new._kfunc_parent = self
            if var or given:
                return super(cls, new).__call__(*var, **given)
            return new

        @property
        def params(self):
            """Parameters stored in the instance (read-only): constructor
            defaults, overridden by the values given; for SDE classes the
            SDE-specific ``args`` as well."""
            out = {k: v for k, v in self._kfunc_init_args.items()
                   if k not in _EXTENSIONS}
            out.update(self._kfunc_params)
            if is_sde:
                out.update(self.args)
            return out

    wrapper.__new__.__wrapped__ = f.__init__
    wrapper.__call__.__wrapped__ = f.__call__
    _named_like(f.__call__, wrapper.__call__)
    setattr(wrapper, _MARK, True)
    return _named_like(f, wrapper)


def _wrap_function(nvar):
    def decorator(f):
        if inspect.isclass(f):
            raise SyntaxError('improper use of the kfunc decorator: '
                              '@kfunc(nvar=k) applies to functions, @kfunc to '
                              'classes')
        sig = [(k, p) for k, p in inspect.signature(f).parameters.items()
               if p.kind != p.VAR_KEYWORD]
        if not 0 < nvar <= len(sig):
            raise ValueError('expecting 0 < nvar <= {}, not {}'
                             .format(len(sig), nvar))
        if any(p.kind != p.KEYWORD_ONLY for _, p in sig[nvar:]):
            raise TypeError('cannot wrap {} as a kfunc: after its nvar={} '
                            'variables, its parameters should be keyword-only'
                            .format(f, nvar))

        class function_kfunc:
            def __init__(self, **params):
                self._kfunc_params = params

            def __call__(self, *var, **given):
                return f(*var, **given, **self._kfunc_params)

        _named_like(f, function_kfunc)
        wrapped = _wrap_class(function_kfunc)
        # the function's own split of variables and parameters (the generic
        # __init__(**params) / __call__(*var, **given) signatures say nothing)
        wrapped._kfunc_init_args = {k: p.default for k, p in sig[nvar:]}
        wrapped._kfunc_call_args = dict(sig[:nvar])
        wrapped.__wrapped__ = f
        return wrapped
    return decorator


def kfunc(f=None, *, nvar=None):
    """Wrap a class (``kfunc(cls)`` / ``@kfunc``) or a function
    (``@kfunc(nvar=k)``) as a kfunc -- see the module docstring."""
    on_class = inspect.isclass(f) and nvar is None
    on_function = f is None and nvar is not None
    if not (on_class or on_function):
        raise SyntaxError('improper use of the kfunc decorator: @kfunc applies '
                          'to classes, @kfunc(nvar=k) to functions')
    return _wrap_class(f) if on_class else _wrap_function(nvar)


def iskfunc(cls_or_object):
    """True for kfunc classes and their instances (reference kfun.py:358)."""
    return hasattr(cls_or_object, _MARK)

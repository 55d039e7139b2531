"""Wire compatibility with nufhe's pickle-based serialization (reference: nufhe/lwe.py:207-243,
api_low_level.py:116-148,198-232, bootstrap.py:78-86, tgsw.py:116-124, tlwe.py:135-145, polynomials.py:72-80).

A nufhe dump is a sequence of pickles: parameter objects (`nufhe.api_low_level.NuFHEParameters`,
`nufhe.lwe.LweParams`, `nufhe.tlwe.TLweParams`, `nufhe.tgsw.TGswParams`) interleaved with NumPy arrays.  The
parameter classes here have the same attribute names, so the only thing that differs is the module path
recorded in the pickle.

    install_nufhe_aliases()        -> dumps made by the reference load here (`nufhe.*` resolves to this package)
    use_reference_pickle_paths()   -> dumps made here also carry `nufhe.*` paths, i.e. the reference can load them
"""
import importlib
import sys

_MODULES = ['api_low_level', 'lwe', 'tlwe', 'tgsw', 'bootstrap', 'polynomials', 'gates', 'performance',
            'random_numbers', 'numeric_functions', 'api_high_level', 'operators_integer']
_PARAM_CLASSES = [('api_low_level', 'NuFHEParameters'), ('lwe', 'LweParams'), ('tlwe', 'TLweParams'),
                  ('tgsw', 'TGswParams')]


def install_nufhe_aliases():
    """Make `import nufhe...` / pickled `nufhe.*` class paths resolve to nufhe_b200 (only if the real nufhe
    is not already imported)."""
    pkg = importlib.import_module('nufhe_b200')
    if 'nufhe' in sys.modules and sys.modules['nufhe'] is not pkg:
        raise RuntimeError('a different `nufhe` package is already imported')
    sys.modules['nufhe'] = pkg
    for name in _MODULES:
        sys.modules['nufhe.' + name] = importlib.import_module('nufhe_b200.' + name)


def use_reference_pickle_paths():
    """Record `nufhe.<module>` as the module of the four pickled parameter classes, so that files dumped here
    can be loaded by the reference as well (and vice versa, via install_nufhe_aliases)."""
    install_nufhe_aliases()
    for mod, cls in _PARAM_CLASSES:
        getattr(importlib.import_module('nufhe_b200.' + mod), cls).__module__ = 'nufhe.' + mod

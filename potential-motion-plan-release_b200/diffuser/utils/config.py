"""Pickle-compatible stand-in for the reference's Config (diffuser/utils/config.py:22-79): a
class object + constructor kwargs.  Reference pickles name this module path and the dotted
paths of the model classes, which is why this mirror keeps both."""
import collections.abc
import importlib
import os
import pickle


def import_class(_class):
    if not isinstance(_class, str):
        return _class
    repo = __name__.split('.')[0]
    mod, _, name = _class.rpartition('.')
    return getattr(importlib.import_module('%s.%s' % (repo, mod)), name)


class Config(collections.abc.Mapping):
    def __init__(self, _class, verbose=False, savepath=None, device=None, **kwargs):
        self._class = import_class(_class)
        self._device = device
        self._dict = dict(kwargs)
        if savepath is not None:
            savepath = os.path.join(*savepath) if isinstance(savepath, tuple) else savepath
            self.savepath = savepath
            with open(savepath, 'wb') as f:
                pickle.dump(self, f)

    def __repr__(self):
        return 'Config(%s, %s)' % (self._class, sorted(self._dict))

    def __iter__(self):
        return iter(self._dict)

    def __getitem__(self, item):
        return self._dict[item]

    def __len__(self):
        return len(self._dict)

    def __getattr__(self, attr):
        if attr == '_dict':
            self.__dict__['_dict'] = {}
            return self.__dict__['_dict']
        try:
            return self._dict[attr]
        except KeyError:
            raise AttributeError(attr)

    def __call__(self, *args, **kwargs):
        inst = self._class(*args, **kwargs, **self._dict)
        return inst.to(self._device) if self._device else inst

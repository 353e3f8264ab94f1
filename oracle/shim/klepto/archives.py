import os
import pickle


class file_archive(dict):
    def __init__(self, name, dict=None, **kw):
        super().__init__()
        self._name = name
        if dict:
            self.update(dict)

    def dump(self):
        with open(self._name, 'wb') as f:
            pickle.dump(builtins_dict(self), f, protocol=pickle.HIGHEST_PROTOCOL)

    def load(self):
        if os.path.isfile(self._name):
            with open(self._name, 'rb') as f:
                self.update(pickle.load(f))


def builtins_dict(d):
    return {k: v for k, v in d.items()}

"""Model factory registry base (mirror of /root/reference/src/margipose/model_factory.py:5-18).

The reference matches versions with the third-party `semantic_version` package (requirements
pin 2.6.0), which is not a dependency here: `Version` / `Spec` restate the two behaviours the
registry uses -- parsing 'MAJOR.MINOR.PATCH' and caret ranges ('^6.0.0' accepts >=6.0.0, <7.0.0).
"""
from abc import ABC, abstractmethod


class Version:
    def __init__(self, text):
        core = str(text).split('-')[0].split('+')[0]
        parts = core.split('.')
        if len(parts) != 3 or not all(p.isdigit() for p in parts):
            raise ValueError('Invalid version string: %r' % (text,))
        self.major, self.minor, self.patch = (int(p) for p in parts)

    def _key(self):
        return (self.major, self.minor, self.patch)

    def __eq__(self, other):
        return isinstance(other, Version) and self._key() == other._key()

    def __lt__(self, other):
        return self._key() < other._key()

    def __le__(self, other):
        return self._key() <= other._key()

    def __hash__(self):
        return hash(self._key())

    def __str__(self):
        return '%d.%d.%d' % self._key()


class Spec:
    """Caret requirement specs ('^X.Y.Z'), the only form the reference's factories use."""

    def __init__(self, text):
        if not str(text).startswith('^'):
            raise ValueError('only caret version specs are supported: %r' % (text,))
        self.base = Version(str(text)[1:])
        b = self.base
        if b.major > 0:
            self.upper = (b.major + 1, 0, 0)
        elif b.minor > 0:
            self.upper = (0, b.minor + 1, 0)
        else:
            self.upper = (0, 0, b.patch + 1)

    def match(self, version):
        if not isinstance(version, Version):
            version = Version(version)
        return self.base._key() <= version._key() < self.upper

    def __contains__(self, version):
        return self.match(version)


class ModelFactory(ABC):
    def __init__(self, model_type, version_spec):
        super().__init__()
        self.model_type = model_type
        self.version_spec = Spec(version_spec)

    def is_for(self, model_type, version):
        """Check if this factory is responsible for the given model type and version."""
        return model_type == self.model_type and version in self.version_spec

    @abstractmethod
    def create(self, model_desc):
        assert self.is_for(model_desc['type'], model_desc['version']), \
            'model_desc does not match this factory'

from . import adam

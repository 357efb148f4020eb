"""B200-native implementation of the baseband-tasks dedispersion and
channelization hot path, behind the baseband-tasks Task / FFTMaker API."""
__version__ = '0.1'

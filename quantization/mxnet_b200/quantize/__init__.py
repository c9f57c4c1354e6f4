#-*- coding: utf-8 -*-
from . import convert

from . import initialize

from . import freeze

from . import distribution_calibrate

from .utils import *

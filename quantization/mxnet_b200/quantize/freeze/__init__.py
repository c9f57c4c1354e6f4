#-*- coding: utf-8 -*-
from .merge_bn import *
from .freeze import *

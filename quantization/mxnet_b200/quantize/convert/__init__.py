#-*- coding: utf-8
from .convert_conv2d import *

from .convert_act import *

from .convert_bn import *

from .convert import *

from .convert_dense import *

from .ste_func import LinearQuantizeSTE

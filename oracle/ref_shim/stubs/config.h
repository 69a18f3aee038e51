/* stand-in for the generated config.h: nothing the pixel code needs */
#pragma once

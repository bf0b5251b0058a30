// Included by the reference sources, nothing from it is used on this path.
#pragma once

// TEST INFRASTRUCTURE ONLY (oracle/_ref build): included by GroundGrid.cpp:39, unused.
#pragma once

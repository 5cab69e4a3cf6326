// TEST INFRASTRUCTURE ONLY (oracle/_ref build): included by GroundGrid.cpp:38, unused.
#pragma once

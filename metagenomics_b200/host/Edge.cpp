// Edge.cpp -- see Edge.h (MetaGenomics/Edge.cpp:17-131).
#include "Edge.h"

Edge::Edge(void)
	: source(NULL), destination(NULL), overlapOrientation(0), overlapOffset(0), listOfReads(new vector<UINT64>),
	  listOfOverlapOffsets(new vector<UINT16>), listOfOrientations(new vector<UINT8>), reverseEdge(NULL), transitiveRemovalFlag(false),
	  flow(0), coverageDepth(0), SD(0)
{
}

Edge::Edge(Read *from, Read *to, UINT64 orient, UINT64 length)
	: listOfReads(NULL), listOfOverlapOffsets(NULL), listOfOrientations(NULL), reverseEdge(NULL), SD(0)
{
	makeEdge(from, to, orient, length);
}

Edge::Edge(Read *from, Read *to, UINT64 orient, UINT64 length, vector<UINT64> *listReads, vector<UINT16> *listOverlapOffsets, vector<UINT8> *listOrientations)
	: listOfReads(NULL), listOfOverlapOffsets(NULL), listOfOrientations(NULL), reverseEdge(NULL), SD(0)
{
	makeEdge(from, to, orient, length, listReads, listOverlapOffsets, listOrientations);
}

Edge::~Edge()
{
	delete listOfReads;
	delete listOfOverlapOffsets;
	delete listOfOrientations;
}

bool Edge::makeEdge(Read *from, Read *to, UINT64 orient, UINT64 length)
{
	return makeEdge(from, to, orient, length, new vector<UINT64>, new vector<UINT16>, new vector<UINT8>);
}

bool Edge::makeEdge(Read *from, Read *to, UINT64 orient, UINT64 length, vector<UINT64> *listReads, vector<UINT16> *listOverlapOffsets, vector<UINT8> *listOrientations)
{
	source = from;
	destination = to;
	overlapOrientation = (UINT8)orient;
	overlapOffset = length;
	transitiveRemovalFlag = false;
	flow = 0;
	coverageDepth = 0;
	listOfReads = listReads;
	listOfOverlapOffsets = listOverlapOffsets;
	listOfOrientations = listOrientations;
	return true;
}

bool Edge::setReverseEdge(Edge *edge)
{
	reverseEdge = edge;
	return true;
}

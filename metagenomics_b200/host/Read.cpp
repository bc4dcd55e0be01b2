// Read.cpp -- see Read.h. addMatePair follows MetaGenomics/Read.cpp:132-166.
#include "Read.h"
#include "Dataset.h"

Read::Read(void)
	: readNumber(0), frequency(0), length(0), owner(NULL), coverageDepth(0), locationInDataset(0), isContainedRead(false), superReadID(0)
{
	matePairList = new vector<MPlist>;
	listOfEdgesForward = new vector<Edge *>;
	locationOnEdgeForward = new vector<UINT64>;
	listOfEdgesReverse = new vector<Edge *>;
	locationOnEdgeReverse = new vector<UINT64>;
}

Read::~Read(void)
{
	delete matePairList;
	delete listOfEdgesForward;
	delete locationOnEdgeForward;
	delete listOfEdgesReverse;
	delete locationOnEdgeReverse;
}

bool Read::setReadNumber(UINT64 id)
{
	if (id < 1) throw OgbFailure(OGB_E_ARG, "ID less than 1.");
	readNumber = id;
	return true;
}

bool Read::setFrequency(UINT32 freq)
{
	if (freq < 1) throw OgbFailure(OGB_E_ARG, "Frequency less than 1.");
	frequency = freq;
	return true;
}

string Read::getStringForward(void) const { return owner->readString(readNumber, 0); }
string Read::getStringReverse(void) const { return owner->readString(readNumber, 1); }

// A mate pair is stored once per (mate, orientation, dataset): Read.cpp:132-166 skips exact repeats.
bool Read::addMatePair(Read *r, UINT8 orientation, UINT64 datasetNumber)
{
	UINT64 id = r->getReadNumber();
	for (size_t i = 0; i < matePairList->size(); i++)
		if (matePairList->at(i).matePairID == id && matePairList->at(i).matePairOrientation == orientation && matePairList->at(i).datasetNumber == datasetNumber)
			return true;
	MPlist m;
	m.matePairID = id;
	m.matePairOrientation = orientation;
	m.datasetNumber = (UINT8)datasetNumber;
	matePairList->push_back(m);
	return true;
}
